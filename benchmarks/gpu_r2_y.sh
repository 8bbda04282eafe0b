#!/bin/bash
# Round 2, GPU visit Y (1 GPU): persistent one-wave gather (lookup_flat), rows per bulk copy in the forward,
# backward variant thresholds, bench with the candidate defaults.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02y}
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "register_variants or rows_per_bulk or fused_launch or interaction_forward_one" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
timeout 300 python benchmarks/ab_bwd.py --variants 1 2 3 --B 2368 3000 4096 8192 > $O/${TAG}_ab_bwd.jsonl 2> $O/${TAG}_ab_bwd.err; echo "ab bwd rc=$?"; cut -c1-140 $O/${TAG}_ab_bwd.jsonl
for r in 1 2 4; do
  timeout 200 python benchmarks/hotpath.py --workload terabyte --B 2048 --only interaction_fwd --nb 16 --small-tables --opt fwd_rows_per_copy=$r > $O/${TAG}_fwd_rpc$r.json 2>> $O/${TAG}_hot.err
  python -c "import json;r=json.load(open('$O/${TAG}_fwd_rpc$r.json'));print('fwd rows_per_copy $r', round(r['interaction_fwd']['us'],2), round(r['interaction_fwd']['frac_hbm'],3))"
done
for f in 2 1; do
  timeout 200 python benchmarks/cta_timeline.py --opt lookup_flat=$f --opt fwd_rows_per_copy=$((f == 1 ? 2 : 1)) > $O/${TAG}_cta_timeline_flat$f.jsonl 2>> $O/${TAG}_cta.err
  grep -E '"lookup_sort"|"interaction_fwd"' $O/${TAG}_cta_timeline_flat$f.jsonl | cut -c1-330
  timeout 200 python benchmarks/hotpath.py --workload terabyte --B 2048 --no-interaction --nb 16 --opt lookup_flat=$f > $O/${TAG}_hot_flat$f.json 2>> $O/${TAG}_hot.err
  python -c "import json;r=json.load(open('$O/${TAG}_hot_flat$f.json'));print('lookup_flat $f', {k:(round(r[k]['us'],2), round(r[k].get('frac_hbm',0),3)) for k in ('lookup','lookup_sort','update_only','embedding_chain')})"
done
timeout 200 python benchmarks/hotpath.py --workload kaggle --B 2048 --no-interaction --nb 16 --opt lookup_flat=1 > $O/${TAG}_hot_kaggle_flat1.json 2>> $O/${TAG}_hot.err
timeout 200 python benchmarks/hotpath.py --workload kaggle --B 2048 --no-interaction --nb 16 --opt lookup_flat=2 > $O/${TAG}_hot_kaggle_flat2.json 2>> $O/${TAG}_hot.err
for f in 1 2; do python -c "import json;r=json.load(open('$O/${TAG}_hot_kaggle_flat$f.json'));print('kaggle lookup_flat $f', {k:(round(r[k]['us'],2), round(r[k].get('frac_hbm',0),3)) for k in ('lookup','lookup_sort','update_only','embedding_chain')})"; done
timeout 300 python bench.py --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_base.json 2> $O/${TAG}_bench_base.err; echo "bench base rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-host-leg --opt lookup_flat=1 > $O/${TAG}_bench_flat.json 2> $O/${TAG}_bench_flat.err; echo "bench flat rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-host-leg --opt lookup_flat=1 --opt fwd_rows_per_copy=2 > $O/${TAG}_bench_flat_rpc2.json 2> $O/${TAG}_bench_flat_rpc2.err; echo "bench flat rpc2 rc=$?"
python - <<PY
import json
for f in ("bench_base","bench_flat","bench_flat_rpc2"):
    try:
        r=json.loads(open("$O/${TAG}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(r['value']), round(r['ms_per_step'],4), round(r['e2e']['value']), r['roofline']['kernel'], round(r['roofline']['frac'],3), {k:round(v,3) if isinstance(v,float) else v for k,v in r['embedding'].items() if k in ('us','frac_hbm','back_to_back_us','frac_hbm_back_to_back')})
        for k,v in r['kernels'].items():
            if not k.startswith('_'): print("   ", k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('in_step_us','back_to_back_us','frac_hbm','frac_hbm_back_to_back')})
    except Exception as e: print(f,"unreadable",e)
PY
