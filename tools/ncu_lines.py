"""Joins an ncu SASS source page with nvdisasm line info: executed warp-instructions and stall
samples per CUDA source line.  usage: python tools/ncu_lines.py report.ncu-rep file.cubin kernel_substr [topN]"""
import csv, re, subprocess, sys
rep, cubin, ksub = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# collect (line) per instruction, in order, for the wanted function
lines = []; cur = None; infn = False; inl = None
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if l.startswith("//--------------------- .text."):
        infn = ksub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ie, sm, so = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
body = [r for r in rows[hi + 1:] if len(r) > ie]
assert len(body) == len(lines), (len(body), len(lines))
agg = {}
tot = 0; tots = 0
for r, ln in zip(body, lines):
    n = int(r[ie]); s = int(r[sm]) if r[sm].isdigit() else 0
    a = agg.setdefault(ln, [0, 0]); a[0] += n; a[1] += s; tot += n; tots += s
src = {}
print(f"total warp-instructions {tot}, samples {tots}")
for ln, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = ""
    if ln:
        try:
            if ln[0] not in src:
                src[ln[0]] = open("/root/repo/cramjam_b200/csrc/" + ln[0]).read().splitlines()
            text = src[ln[0]][ln[1] - 1].strip()[:110]
        except Exception:
            pass
    print(f"{100*n/tot:5.1f}% inst {100*s/max(tots,1):5.1f}% samp  {ln}  {text}")

# optional: aggregate by line ranges given as name:lo-hi,...
if len(sys.argv) > 5:
    ranges = []
    for part in sys.argv[5].split(","):
        name, r = part.split(":"); lo, hi = r.split("-"); ranges.append((name, int(lo), int(hi)))
    acc = {name: [0, 0] for name, _, _ in ranges}; acc["other"] = [0, 0]
    for ln, (n, s) in agg.items():
        key = "other"
        if ln and ln[0] in ("lz_decode.cu", "lz_decode.cuh"):
            for name, lo, hi in ranges:
                if lo <= ln[1] <= hi:
                    key = name; break
        acc[key][0] += n; acc[key][1] += s
    print("--- by region ---")
    for k, (n, s) in acc.items():
        print(f"{k:12s} {100*n/tot:5.1f}% inst  {100*s/max(tots,1):5.1f}% samples")
