#!/bin/bash
# Round 2, GPU visit J (1 GPU): the two-warps-per-sample forward inside the step (device clock), tests.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02j}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward or interaction" > $O/${TAG}_pytest_fwd.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_fwd.log
for B in 2048 4096; do for m in 0 2; do
  timeout 300 python benchmarks/hotpath.py --workload terabyte --B $B --small-tables --only interaction_fwd --nb 16 --opt fwd_ksplit=$m > $O/${TAG}_fwd_mode${m}_B$B.json 2>> $O/hot_j.err
done; done
timeout 600 python bench.py --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_fwd_*.json")):
    try:
        r=json.load(open(f)); print(f.split('/')[-1], round(r['interaction_fwd']['us'],2), round(r['interaction_fwd']['frac_hbm'],3))
    except Exception as e:
        print(f, "unreadable", e)
r=json.load(open("$O/${TAG}_bench_n1.json"))
print("bench", round(r["value"]), round(r["ms_per_step"],4), {k:(round(v.get('back_to_back_us',0),2), round(v['in_step_us'],2), round(v.get('frac_hbm',0),3)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r.get('embedding',{}).get('frac_hbm'), r.get('embedding',{}).get('frac_hbm_back_to_back'))
PY
