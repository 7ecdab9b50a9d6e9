"""Per-kernel SASS instruction counts of the shipped library (evidence that the hot kernels are what DESIGN.md says):
    python tools/sass_counts.py > profiles/r02_sass_counts.txt
DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind), UBLKCP = cp.async.bulk (TMA bulk copy,
non-tensor), STAS = st.async to distributed shared memory, SYNCS = mbarrier operations, LDGSTS = cp.async,
UTMALDG / UTC*MMA = tensor-map TMA / tcgen05 (expected: none -- no FP64 path exists there)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "admmsolver_b200", "libadmm_b200.so")
MNEMONICS = ["DMMA", "DFMA", "UBLKCP", "UTMALDG", "UTCHMMA", "UTCQMMA", "STAS", "SYNCS", "LDGSTS", "ACQBULK", "CCTL", "BAR", "ATOMG", "RED"]
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
names = iter(filt)
counts = None
rows = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        counts = collections.Counter()
        rows.append((next(names), counts))
        continue
    if counts is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts["total"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                counts[mn] += 1
sha = hashlib.sha256(open(lib, "rb").read()).hexdigest()[:16]
src = hashlib.sha256(b"".join(open(os.path.join(ROOT, "admmsolver_b200", "csrc", f), "rb").read()
                              for f in ("common.cuh", "primitives.cu", "spm.cu", "bp.cu", "peer.cu"))).hexdigest()[:16]
print(f"# libadmm_b200.so sha256[:16] = {sha}; csrc (common.cuh primitives.cu spm.cu bp.cu peer.cu) sha256[:16] = {src}")
print("# nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17; cuobjdump -sass; counts per kernel")
print("# " + " ".join(f"{m:>7s}" for m in ["total"] + MNEMONICS) + "  kernel")
short = lambda n: re.sub(r"\(.*", "", n.replace("admm::", "").replace("(anonymous namespace)::", ""))
for name, c in sorted(rows, key=lambda r: -r[1]["DMMA"]):
    if "-a" not in sys.argv and c["DMMA"] == 0 and c["UBLKCP"] == 0 and c["STAS"] == 0:
        continue
    print("  " + " ".join(f"{c[m]:7d}" for m in ["total"] + MNEMONICS) + "  " + short(name))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print("# whole library: " + ", ".join(f"{m} {tot[m]}" for m in MNEMONICS) + f"; {len(rows)} kernels")
