"""Compact JSON summary of an ncu report (the format bench.py's roofline.traffic reads from profiles/):
    python scripts/ncu_raw_to_json.py report.ncu-rep out.json [kernel-name substring]
One record per profiled launch: kernel name, grid / block, duration, DRAM bytes, pipe and issue utilisation, registers,
shared memory, L1 wavefront and bank-conflict counters."""
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(head)}
out = []
for r in rows[2:]:
    if len(r) < len(head) or (len(sys.argv) > 3 and sys.argv[3] not in r[ix["Kernel Name"]]):
        continue
    rec = {"Kernel Name": r[ix["Kernel Name"]], "Block Size": r[ix["Block Size"]], "Grid Size": r[ix["Grid Size"]]}
    for k in KEEP:
        if k in ix:
            rec[f"{k} [{units[ix[k]]}]" if units[ix[k]] else k] = r[ix[k]]
    out.append(rec)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(f"{len(out)} launches -> {sys.argv[2]}")
