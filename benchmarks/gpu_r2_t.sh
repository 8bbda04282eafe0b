#!/bin/bash
# Round 2, GPU visit T (1 GPU): device-clock stamps for grids above 4096 CTAs (atomic min / max into folded slots);
# bench at per-GPU batch 16384 re-measured with them, default bench re-checked.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02t}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_clock or interaction_forward_one or fused_launch" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
timeout 400 python bench.py --batch 16384 --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_terabyte_B16384_n1.json 2> $O/${TAG}_bench_terabyte_B16384_n1.err; echo "B16384 rc=$?"
timeout 400 python bench.py --no-cpu-baseline > $O/${TAG}_bench_terabyte_n1.json 2> $O/${TAG}_bench_terabyte_n1.err; echo "B2048 rc=$?"
python - <<PY
import json
for f in ("bench_terabyte_B16384_n1","bench_terabyte_n1"):
    try:
        r=json.loads(open("$O/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, r.get('value'), r.get('ms_per_step'), (r.get('e2e') or {}).get('value'), r['roofline']['kernel'], r['roofline']['frac'], r['embedding'])
        for k,v in r['kernels'].items():
            print("   ", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('in_step_us','event_pair_us','back_to_back_us','frac_hbm','frac_hbm_back_to_back')})
    except Exception as e:
        print(f, "unreadable", e)
PY
