"""SASS mnemonic counts per kernel of the built library (static instruction counts): which kernels use TMA / mbarriers / 16-bit SIMD.
usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import subprocess, re, collections, os
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chalkydri_b200", "libchalkydri_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UBLKCP", "UTMALDG", "SYNCS", "VIMNMX3", "VIMNMX", "REDUX", "MATCH", "DADD", "DMUL", "LDS", "STS", "SHFL", "ATOMS", "ATOMG"]
cur, cnt = None, collections.defaultdict(collections.Counter)
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        op = m.group(1).split(".")[0]
        cnt[cur]["total"] += 1
        if op in KEYS:
            cnt[cur][op] += 1
names = list(cnt.keys())
dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
print("# SASS mnemonic counts per kernel of chalkydri_b200/libchalkydri_b200.so (cuobjdump -sass, sm_100a); static instruction counts")
print("# UBLKCP = cp.async.bulk (1-D TMA loads; the threshold kernel's bulk stores), UTMALDG = cp.async.bulk.tensor load, SYNCS = mbarrier operations,")
print("# VIMNMX3 / VIMNMX = 16-bit SIMD min/max (threshold), REDUX / MATCH = warp reductions / match.any (CCL, cluster passes)")
print("%-100s %6s " % ("kernel", "total") + " ".join("%7s" % k for k in KEYS))
for k, d in sorted(zip(names, dem), key=lambda x: x[1]):
    c = cnt[k]
    name = re.sub(r"\((?:const |unsigned |cb::|CUtensor).*", "", d)[:100]
    print("%-100s %6d " % (name, c["total"]) + " ".join("%7d" % c[x] for x in KEYS))
