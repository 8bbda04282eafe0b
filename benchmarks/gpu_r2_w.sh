#!/bin/bash
# Round 2, GPU visit W (1 GPU): backward register variants (one wave at 128 registers), shared-memory same-launch
# fix-up of the sparse update; parity, per-CTA timelines, cold back-to-back timings, bench per variant.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02w}
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "register_variants or sparse_sgd or interaction_backward or interaction_warp or validate_golden or update or bf16 or terabyte" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
for v in 0 1 2 3; do
  timeout 200 python benchmarks/cta_timeline.py --opt bwd_variant=$v > $O/${TAG}_cta_timeline_v$v.jsonl 2>> $O/${TAG}_cta.err; echo "timeline v$v rc=$?"
  grep -E '"interaction_bwd"|"update"' $O/${TAG}_cta_timeline_v$v.jsonl | cut -c1-420
  for B in 2048 16384; do
    timeout 200 python benchmarks/hotpath.py --workload terabyte --B $B --only interaction_bwd --nb $((B == 2048 ? 16 : 4)) --opt bwd_variant=$v > $O/${TAG}_bwd_v${v}_B$B.json 2>> $O/${TAG}_hot.err
    python -c "import json;r=json.load(open('$O/${TAG}_bwd_v${v}_B$B.json'));print('bwd v$v B$B', round(r['interaction_bwd']['us'],2), round(r['interaction_bwd']['frac_hbm'],3))"
  done
done
timeout 300 python benchmarks/ab_update.py --B 2048 --cases two > $O/${TAG}_ab_update.jsonl 2> $O/${TAG}_ab_update.err; echo "ab rc=$?"; cat $O/${TAG}_ab_update.jsonl
for v in 0 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-host-leg --opt bwd_variant=$v > $O/${TAG}_bench_v$v.json 2> $O/${TAG}_bench_v$v.err; echo "bench v$v rc=$?"
done
python - <<PY
import json
for f in ("bench_v0","bench_v1","bench_v2"):
    try:
        r=json.loads(open("$O/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(r['value']), round(r['ms_per_step'],4), round(r['e2e']['value']), r['roofline']['kernel'], round(r['roofline']['frac'],3), {k:round(v,3) if isinstance(v,float) else v for k,v in r['embedding'].items() if k in ('us','frac_hbm','back_to_back_us','frac_hbm_back_to_back')})
        for k,v in r['kernels'].items():
            if not k.startswith('_'): print("   ", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('in_step_us','back_to_back_us','frac_hbm','frac_hbm_back_to_back')})
    except Exception as e: print(f,"unreadable",e)
PY
