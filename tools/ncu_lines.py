"""Top source lines of a kernel in an ncu report by instructions executed / stall samples.
  python tools/ncu_lines.py report.ncu-rep kernel_regex [launch_skip] [top_n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = ""
data = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); ti = hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > ci and r[0] not in ("", "Line No"):
        try:
            data.append((int(r[ci]), int(r[si]), int(r[ti]), fname, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print("total warp instructions", tot, "samples", tots)
print("--- by instructions")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]:10d} {100*d[0]/max(tot,1):5.1f}%  smp {100*d[1]/max(tots,1):5.1f}%  thr/inst {d[2]/max(d[0],1):4.1f}  {d[3]}:{d[4]}  {d[5]}")
print("--- by stall samples")
for d in sorted(data, key=lambda x: -x[1])[:top // 2]:
    print(f"{d[0]:10d} {100*d[0]/max(tot,1):5.1f}%  smp {100*d[1]/max(tots,1):5.1f}%  thr/inst {d[2]/max(d[0],1):4.1f}  {d[3]}:{d[4]}  {d[5]}")
