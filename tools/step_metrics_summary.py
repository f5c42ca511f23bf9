"""Per-kernel summary of a tools/step_metrics.sh capture: python tools/step_metrics_summary.py <csv> [sequences per launch]"""
import csv, collections, re, sys, json
path = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lines=[l for l in open(path) if not l.startswith('==')]
per=collections.OrderedDict()
for r in csv.DictReader(lines):
    name=re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","").replace("<unnamed>::","")
    name=re.sub(r"<.*","",name)
    tpl=re.search(r"k_odom_(search|stage)<(\d)>", r["Kernel Name"])
    if tpl: name=f"k_odom_{tpl.group(1)}_{'surf' if tpl.group(2)=='0' else 'corner'}"
    m=r["Metric Name"]; v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    d=per.setdefault(name,collections.defaultdict(list))
    if m=="gpu__time_duration.sum": v = v/1000.0 if u=="ns" else (v*1000.0 if u=="ms" else v)
    if m.startswith("dram__bytes"): v = v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1)
    d[m].append(v)
tot_t=sum(sum(d["gpu__time_duration.sum"]) for d in per.values()); tot_i=sum(sum(d["smsp__inst_executed.sum"]) for d in per.values())
print(f"launches {sum(len(d['gpu__time_duration.sum']) for d in per.values())}  total {tot_t:.0f} us (serialised, cold cache: compare shares)  {tot_i/1e6:.1f} M warp instructions; {B} sequences per launch")
traffic={}
for n,d in sorted(per.items(), key=lambda kv:-sum(kv[1]["gpu__time_duration.sum"])):
    t=d["gpu__time_duration.sum"]; i=d["smsp__inst_executed.sum"]; rd=d["dram__bytes_read.sum"]; wr=d["dram__bytes_write.sum"]
    act=d["smsp__thread_inst_executed_per_inst_executed.ratio"]; occ=d["sm__warps_active.avg.pct_of_peak_sustained_active"]; l2=d["lts__t_sector_hit_rate.pct"]
    print(f"{n:24s} n={len(t):4d} avg_us={sum(t)/len(t):8.1f} time%={100*sum(t)/tot_t:5.1f} inst%={100*sum(i)/tot_i:5.1f} Minst/launch={sum(i)/len(i)/1e6:7.2f} dramMB/launch={(sum(rd)+sum(wr))/len(t)/1e6:7.2f} act_thr={sum(act)/len(act):5.1f} warps%={sum(occ)/len(occ):5.1f} L2hit%={sum(l2)/len(l2):5.1f}")
    traffic[n]={"dram_bytes_per_sequence": (sum(rd)+sum(wr))/len(t)/B, "launches": len(t)}
if len(sys.argv) > 3:
    json.dump(traffic, open(sys.argv[3],"w"), indent=1)
