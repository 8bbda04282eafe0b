#!/bin/bash
# Round 2, GPU visit U (1 GPU): parity of the changed kernels (forward tail pairs as FP32 dot products, L2
# prefetch hints in the sparse update), A/B of the update switches, bench with and without the prefetch.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02u}
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prefetch or interaction_forward or interaction_warp or known_answer or validate_golden or inline_fixup or any_tile" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
timeout 300 python benchmarks/ab_update.py --B 2048 16384 > $O/${TAG}_ab_update.jsonl 2> $O/${TAG}_ab_update.err; echo "ab rc=$?"; cat $O/${TAG}_ab_update.jsonl
timeout 200 python benchmarks/hotpath.py --workload terabyte --B 2048 --only interaction_fwd --nb 16 > $O/${TAG}_fwd_B2048.json 2>> $O/${TAG}_hot.err
timeout 200 python benchmarks/hotpath.py --workload terabyte --B 16384 --only interaction_fwd --nb 4 > $O/${TAG}_fwd_B16384.json 2>> $O/${TAG}_hot.err
timeout 300 python bench.py --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_pf0.json 2> $O/${TAG}_bench_pf0.err; echo "bench pf0 rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-host-leg --opt update_prefetch=3 > $O/${TAG}_bench_pf3.json 2> $O/${TAG}_bench_pf3.err; echo "bench pf3 rc=$?"
python - <<PY
import json
for f in ("fwd_B2048","fwd_B16384"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f)); print(f, round(r['interaction_fwd']['us'],2), round(r['interaction_fwd']['frac_hbm'],3))
    except Exception as e: print(f,"unreadable",e)
for f in ("bench_pf0","bench_pf3"):
    try:
        r=json.loads(open("$O/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(r['value']), round(r['ms_per_step'],4), round(r['e2e']['value']), r['roofline']['kernel'], round(r['roofline']['frac'],3), {k:round(v,3) if isinstance(v,float) else v for k,v in r['embedding'].items() if k in ('us','frac_hbm','back_to_back_us','frac_hbm_back_to_back')})
        for k,v in r['kernels'].items():
            if not k.startswith('_'): print("   ", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('in_step_us','back_to_back_us','frac_hbm','frac_hbm_back_to_back')})
    except Exception as e: print(f,"unreadable",e)
PY
