#!/bin/bash
# Round 2, GPU visit Z (1 GPU): output-stationary streaming backward (4 = register ring, 5 = cp.async ring in shared
# memory) against the resident-T variants; parity, cold A/B, per-CTA timelines, bench in step.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02z}
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "register_variants or interaction_warp or interaction_backward" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
timeout 300 python benchmarks/ab_bwd.py --variants 1 2 3 4 5 --B 2048 4096 16384 > $O/${TAG}_ab_bwd.jsonl 2> $O/${TAG}_ab_bwd.err; echo "ab bwd rc=$?"; cut -c1-140 $O/${TAG}_ab_bwd.jsonl
for v in 4 5; do
  timeout 200 python benchmarks/cta_timeline.py --opt bwd_variant=$v > $O/${TAG}_cta_timeline_v$v.jsonl 2>> $O/${TAG}_cta.err
  grep -E '"interaction_bwd"' $O/${TAG}_cta_timeline_v$v.jsonl | cut -c1-330
done
for v in 0 4 5; do
timeout 300 python bench.py --no-cpu-baseline --no-host-leg --opt bwd_variant=$v > $O/${TAG}_bench_v$v.json 2> $O/${TAG}_bench_v$v.err; echo "bench v$v rc=$?"
done
python - <<PY
import json
for f in ("bench_v0","bench_v4","bench_v5"):
    try:
        r=json.loads(open("$O/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(r['value']), round(r['ms_per_step'],4), round(r['e2e']['value']), r['roofline']['kernel'], round(r['roofline']['frac'],3), {k:round(v,3) if isinstance(v,float) else v for k,v in r['embedding'].items() if k in ('us','frac_hbm','back_to_back_us','frac_hbm_back_to_back')})
        for k,v in r['kernels'].items():
            if k in ('interaction_bwd','update','lookup','interaction_fwd'): print("   ", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('in_step_us','back_to_back_us','frac_hbm','frac_hbm_back_to_back')})
    except Exception as e: print(f,"unreadable",e)
PY
