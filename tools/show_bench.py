"""Pretty-print a bench.py JSON line: headline numbers + per-kernel time per step."""
import json, sys
d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} scans/s  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.3f} ms/step)  launches {d['gpu_launches']}")
r = d['roofline']
print(f"roofline kernel {r['kernel']} avg {r['avg_launch_us']:.1f} us achieved {r['achieved']:.1f} GB/s frac {r['frac']:.4f}")
for k, v in r['kernel_time_share_profiling_pass'].items():
    print(f"  {k:26s} {v:.3f}  ~{v * d['ms_per_step'] * 1000:7.1f} us/step")
print(d['stats'])
if 'cpu_baseline' in d: print(d['cpu_baseline'])
if 'kernel_rooflines' in d:
    kr = d['kernel_rooflines']
    print("kernels timed alone (%s sequences per launch): psf_mean_frac %s" % (kr.get('sequences_per_launch'), kr.get('psf_mean_frac')))
    for k, v in kr.items():
        if isinstance(v, dict):
            print(f"  {k:26s} {v['avg_us']:8.1f} us  {v['GBps']:8.1f} GB/s  frac {v['frac']}")
if d.get('latency_single_sequence'): print(d['latency_single_sequence'])
