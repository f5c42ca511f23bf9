"""Print value / e2e / e2e_alt of bench JSON files."""
import json, sys
for n in sys.argv[1:]:
    try:
        d = json.loads([x for x in open(n) if x.strip().startswith("{")][-1])
        e, a = d.get("e2e", {}), d.get("e2e_alt", {})
        print(f"{n}: value {d['value']:.0f} ({d['ms_per_step']:.3f} ms)  e2e[{e.get('host_points')}] {e.get('value', 0):.0f} ({e.get('ms_per_step', 0):.3f} ms, {e.get('h2d_bytes_per_step', 0) / 1e6:.0f} MB)  "
              f"alt[{a.get('host_points')}] {a.get('value', 0):.0f} ({a.get('ms_per_step', 0):.3f} ms)  roof {d['roofline']['kernel']} {d['roofline']['frac']:.4f}  clocks {d['clocks']}")
    except Exception as ex:
        print(n, "ERR", ex)
