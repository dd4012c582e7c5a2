"""Per-kernel counts of the Blackwell-specific SASS opcodes in libfluxgnn.so (evidence that tcgen05 / TMEM / TMA are
really in the binary): python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gnn_plasma_flux_b200", "libfluxgnn.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FMUL2", "FADD2",
         "FFMA", "DFMA", "HFMA2", "FHFMA", "LDS", "STS", "SHFL", "ELECT"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
name, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"fluxgnn::(\(anonymous namespace\)::)?", "", name)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        counts[name][m.group(1)] += 1
        counts[name]["(all)"] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB)}: static instruction counts per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA),")
print("# UTMALDG/UTMASTG = cp.async.bulk.tensor (tensor-map TMA), SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed fp32x2")
total = collections.Counter()
for k, c in counts.items():
    total.update(c)
    seen = "  ".join(f"{op}={c[op]}" for op in WATCH if c[op])
    print(f"{k[:110]:110s} instr={c['(all)']:6d}  {seen}")
print("TOTAL  " + "  ".join(f"{op}={total[op]}" for op in WATCH))
