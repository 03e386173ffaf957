"""Join an ncu SASS source page (ncu -i rep --page source --csv > src.csv) with nvdisasm -g line info of the cubin.
usage: python tools/sass_profile.py src.csv dis_all.txt <mangled kernel substring> [topN]"""
import collections, csv, re, sys
src, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
cur = None; seq = []
for ln in lines[start + 1:]:
    if ln.startswith("//---------------------") or ln.startswith(".section") or ln.startswith("\t.section"):
        if seq: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: seq.append((cur, m.group(2)))
rows = list(csv.reader(open(src))); hdr = rows[1]; data = rows[2:]
ci = hdr.index("Instructions Executed"); ct = hdr.index("Thread Instructions Executed")
cs = hdr.index("# Samples")
print("sass instrs: profile", len(data), "disasm", len(seq))
n = min(len(data), len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0])
for r, (loc, txt) in zip(data[:n], seq[:n]):
    k = loc or ("?", 0)
    agg[k][0] += int(r[ci]); agg[k][1] += int(r[ct]); agg[k][2] += int(r[cs] or 0)
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("total warp inst", tot, " thread inst", sum(v[1] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.1f}% inst  {100*v[2]/max(ts,1):5.1f}% samples  thr {v[1]/max(v[0],1):5.1f}  {k[0]}:{k[1]}")
