"""Per-CUDA-source-line instruction and stall-sample shares from an ncu report captured with --import-source on:
  ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > lines.csv ; python tools/ncu_lines.py lines.csv [top]"""
import collections, csv, os, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
fn, fpath, hdr, cur = None, None, None, None
per = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = os.path.basename(r[1]); continue
    if len(r) == 2 and r[0] == "Function Name":
        fn = re.sub(r"\(.*", "", re.sub(r".*::", "", r[1])); continue
    if r and r[0] == "Line No":
        hdr = r; isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != "":
        cur = (fpath, int(r[0]), r[1].strip()); continue
    if r[2] in ("...", "") or cur is None:
        continue
    try:
        s, e = int(r[isamp]), int(r[iex])
    except ValueError:
        continue
    d = per.setdefault(fn, collections.OrderedDict())
    a = d.setdefault(cur, [0, 0]); a[0] += s; a[1] += e
for k, d in per.items():
    ts, ti = sum(v[0] for v in d.values()), sum(v[1] for v in d.values())
    print(f"===== {k}: {ts} samples, {ti} warp instructions")
    for (f, ln, src), (s, e) in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {f}:{ln:<4d} inst {100 * e / max(1, ti):5.1f}%  samples {100 * s / max(1, ts):5.1f}%  {src[:96]}")
