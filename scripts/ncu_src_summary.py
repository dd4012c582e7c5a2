"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest
instructions per reason.  usage: python scripts/ncu_src_summary.py file.csv [kernel-index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = b["hdr"]
ix = {n: i for i, n in enumerate(h)}
val = lambda r, k: int(float(r[ix[k]] or 0))
tot = sum(val(r, "# Samples") for r in b["rows"])
print(b["name"], len(b["rows"]), "instructions, total samples", tot)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(val(r, s) for r in b["rows"]) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for key in ("stall_long_sb", "stall_dispatch", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_math", "stall_barrier"):
    top = sorted(b["rows"], key=lambda r: -val(r, key))[:6]
    print("==", key)
    for r in top:
        print(val(r, key), r[ix["Address"]], r[ix["Source"]][:100])
