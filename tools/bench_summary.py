import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.0f scans/s  ms/step %.3f  e2e %.0f scans/s (%.3f ms/step)  launches/step %.1f" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stats"]["launches_per_step"]))
r = d["roofline"]
print("roofline kernel %s avg %.1f us achieved %.1f GB/s frac %.4f" % (r["kernel"], r["avg_launch_us"], r["achieved"], r["frac"]))
tot = d["ms_per_step"] * 1000
for k, v in r["kernel_time_share_profiling_pass"].items():
    print("  %-24s %6.3f  ~%7.1f us/step" % (k, v, v * tot))
print("stats", d["stats"])
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
