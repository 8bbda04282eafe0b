#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01t}
timeout 200 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench_terabyte_n1.json 2> $O/${TAG}_bench_terabyte_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --workload kaggle --no-cpu-baseline > $O/${TAG}_bench_kaggle_n1.json 2> $O/${TAG}_bench_kaggle_n1.err; echo "kaggle rc=$?"
python - <<PY
import json
for f in ("bench_terabyte_n1","bench_kaggle_n1"):
    r=json.load(open("$O/${TAG}_%s.json"%f))
    print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v['avg_us'],2), round(v.get('frac_hbm',0),3)) for k,v in r.get('kernels',{}).items()}, r['roofline']['kernel'], round(r['roofline']['frac'],3), r.get("cpu_baseline",{}).get("value"), r["clocks"]["samples"])
PY
