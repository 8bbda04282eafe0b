#!/bin/bash
# Round 2, final single-GPU visit: smoke, the whole GPU suite, bench (both arms), Kaggle-shaped microbenchmark.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02z}
timeout 300 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py > $O/${TAG}_bench_terabyte_n1.json 2> $O/${TAG}_bench_terabyte_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference_cpu.json 2> $O/${TAG}_bench_reference_cpu.err; echo "bench ref rc=$?"
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 --no-interaction > $O/${TAG}_hot_kaggle_B2048.json 2>> $O/hot_z.err
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 > $O/${TAG}_hot_terabyte_B2048.json 2>> $O/hot_z.err
python - <<PY
import json
for f in ("hot_kaggle_B2048","hot_terabyte_B2048"):
    r=json.load(open("$O/${TAG}_%s.json"%f))
    print(f,{k:(round(r[k]['us'],2), round(r[k].get('frac_hbm',0),3)) for k in ('lookup','sort','lookup_sort','update_only','embedding_chain','interaction_fwd','interaction_bwd') if k in r})
for f in ("bench_terabyte_n1","bench_reference_cpu"):
    r=json.load(open("$O/${TAG}_%s.json"%f))
    print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v.get('back_to_back_us',0),2), round(v['in_step_us'],2), round(v.get('frac_hbm',0),3)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r.get('roofline',{}).get('kernel'), r.get('roofline',{}).get('frac'), r.get('embedding',{}).get('frac_hbm'), r.get('embedding',{}).get('frac_hbm_back_to_back'), r.get("cpu_baseline"), r.get("e2e_host"), r.get("clocks"))
PY
