"""Per-source-line profile from an ncu report captured with --import-source on (kernel built with -lineinfo).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/line_profile.py src.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
def num(s):
    try: return int(s)
    except ValueError: return 0
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        ii = hdr.index("Instructions Executed"); ss = hdr.index("Warp Stall Sampling (All Samples)"); at = hdr.index("Avg. Threads Executed")
        out.append((cur, int(r[0]), r[1].strip()[:90], num(r[ss]), num(r[ii]), r[at]))
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print("stall samples", ts, "warp instructions", ti)
byfile = {}
for o in out: byfile.setdefault(o[0], [0, 0]); byfile[o[0]][0] += o[3]; byfile[o[0]][1] += o[4]
print({k: (round(100 * v[0] / ts, 1), round(100 * v[1] / ti, 1)) for k, v in byfile.items()})
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{100*o[3]/ts:5.1f}% smp {100*o[4]/ti:5.1f}% inst thr {o[5]:>3} {o[0]}:{o[1]}  {o[2]}")
