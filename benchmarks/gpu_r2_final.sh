#!/bin/bash
# Round 2, last GPU visit (1 GPU): the whole GPU suite, then bench lines with the final defaults.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02f}
timeout 170 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest_gpu.log
timeout 60 python bench.py --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_terabyte_n1.json 2> $O/${TAG}_bench_terabyte_n1.err; echo "bench rc=$?"
timeout 40 python bench.py --workload kaggle --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_kaggle_n1.json 2> $O/${TAG}_bench_kaggle_n1.err; echo "kaggle rc=$?"
timeout 60 python bench.py --batch 16384 --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_terabyte_B16384_n1.json 2> $O/${TAG}_bench_terabyte_B16384_n1.err; echo "B16384 rc=$?"
python - <<PY
import json
for f in ("bench_terabyte_n1","bench_kaggle_n1","bench_terabyte_B16384_n1"):
    try:
        r=json.loads(open("$O/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(r['value']), round(r['ms_per_step'],4), round(r['e2e']['value']), r['roofline']['kernel'], round(r['roofline']['frac'],3), {k:round(v,3) if isinstance(v,float) else v for k,v in r['embedding'].items() if k in ('us','frac_hbm','back_to_back_us','frac_hbm_back_to_back')}, r.get('clocks'))
        for k,v in r['kernels'].items():
            if not k.startswith('_'): print("   ", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('in_step_us','back_to_back_us','frac_hbm','frac_hbm_back_to_back')})
    except Exception as e: print(f,"unreadable",e)
PY
