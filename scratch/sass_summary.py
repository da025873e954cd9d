"""Per-kernel SASS mnemonic counts of the built library (evidence for profiles/): python scratch/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "spgemm_b200/libtilespgemm_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, ops = None, collections.defaultdict(collections.Counter)
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        ops[cur][m.group(1).split(".")[0]] += 1
names = subprocess.run(["c++filt"], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
dem = {k: v.split("(")[0].replace("tsg::", "").replace("void ", "") for k, v in zip(ops, names)}
keys = ["UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMMA", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "REDG", "POPC", "SHFL", "VOTE", "MATCH", "BAR"]
print(f"# SASS mnemonic counts per kernel of {so} (cuobjdump -sass, sm_100a)")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, LDGSTS = cp.async, DFMA / DMMA = FP64 FMA / FP64 tensor-core MMA,")
print("# ATOMS / ATOMG / REDG = shared / global atomics, SHFL / VOTE / MATCH = warp collectives")
print("%-44s %6s " % ("kernel", "instrs") + " ".join("%6s" % k for k in keys))
for f, c in sorted(ops.items(), key=lambda x: dem[x[0]]):
    print("%-44s %6d " % (dem[f][:44], sum(c.values())) + " ".join("%6d" % c.get(k, 0) for k in keys))
