#!/bin/bash
# A/B of alternative builds of the library (LEGO_LOAM_B200_LIB): tools/ab_bench.sh <tag> <variant suffix> ...   ("" = the default build)
tag=$1; shift
for v in "$@"; do
  lib=lego_loam_bor_b200/liblego_loam_b200${v:+_$v}.so
  LEGO_LOAM_B200_LIB=$PWD/$lib timeout 300 python bench.py --skip-cpu-baseline --skip-latency > gpurun_out/${tag}_ab_${v:-base}.json 2> gpurun_out/${tag}_ab_${v:-base}.err
  python - "$v" gpurun_out/${tag}_ab_${v:-base}.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][0])
kr = d["kernel_rooflines"]
pick = ["k_odom_search_corner", "k_odom_search_surf", "k_map_knn", "k_map_knn_reuse", "k_feature_ring", "k_odom_stage_corner"]
print("%-8s value %.0f e2e %.0f | " % (sys.argv[1] or "base", d["value"], d["e2e"]["value"]) + "  ".join("%s %.1f" % (k[2:], kr[k]["avg_us"]) for k in pick if k in kr))
PY
done
