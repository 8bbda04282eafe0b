#!/bin/bash
# Round 2, GPU visit S (1 GPU): bench lines of the other SURVEY 8(d) configurations -- config 3 (Kaggle-shaped,
# both arms) and config 4 at a per-GPU batch of 16384.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02s}
timeout 400 python bench.py --workload kaggle > $O/${TAG}_bench_kaggle_n1.json 2> $O/${TAG}_bench_kaggle_n1.err; echo "kaggle rc=$?"
timeout 400 python bench.py --workload kaggle --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_kaggle_reference_cpu.json 2> $O/${TAG}_bench_kaggle_reference_cpu.err; echo "kaggle ref rc=$?"
timeout 400 python bench.py --batch 16384 --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_terabyte_B16384_n1.json 2> $O/${TAG}_bench_terabyte_B16384_n1.err; echo "B16384 rc=$?"
python - <<PY
import json
for f in ("bench_kaggle_n1","bench_kaggle_reference_cpu","bench_terabyte_B16384_n1"):
    try:
        r=json.loads(open("$O/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, r.get('value'), r.get('ms_per_step'), (r.get('e2e') or {}).get('value'), (r.get('roofline') or {}).get('frac'), (r.get('embedding') or {}).get('frac_hbm'), (r.get('embedding') or {}).get('frac_hbm_back_to_back'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $O/${TAG}_bench_kaggle_n1.err
