"""Counts the SASS mnemonics that matter for the evidence (TMA bulk copies, mbarrier, cp.async, vector memory ops) per kernel of
the shipped library.  usage: python tools/sass_summary.py [path/to/libcramjam_cuda.so] > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cramjam_b200", "libcramjam_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = {}
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "STG.E.ENL2.256", "STG.E.128", "LDG.E.128", "LDS.128", "STS.128", "LDS.64", "STS.64", "ATOMS", "ATOMG", "MATCH", "SHFL", "VOTE", "REDUX", "BAR.SYNC", "MEMBAR", "HMMA", "UTMALDG", "UTCHMMA"]
cur = None
counts = collections.OrderedDict()
total = collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[cur][w] += 1
names = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS summary of {os.path.basename(so)} (cuobjdump -sass, sm_100a cubins; instruction counts are static, per kernel)")
print("# UBLKCP = cp.async.bulk (TMA bulk copy; .S.G = global->shared, .G.S = shared->global), SYNCS = mbarrier ops, LDGSTS = cp.async, STG.E.ENL2.256 = 32-byte st.global.v8")
for (k, c), n in zip(counts.items(), names):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("cj::", "")
    parts = [f"{w}={c[w]}" for w in WATCH if c[w]]
    print(f"{n:60s} instructions={c['_all']:5d}  " + " ".join(parts))
