#!/bin/bash
# Round 2, GPU visit B: bucket + rank small sort, grouped inline fix-up, new bench.py.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02b}
timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $O/${TAG}_pytest_gpu.log
for wl in terabyte kaggle; do
  timeout 300 python benchmarks/hotpath.py --workload $wl --B 2048 --no-interaction > $O/${TAG}_hot_${wl}_B2048.json 2>> $O/hot_b.err
done
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 --no-interaction --opt update_two_launches=1 > $O/${TAG}_hot_terabyte_B2048_two_launch.json 2>> $O/hot_b.err
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 4096 --no-interaction > $O/${TAG}_hot_terabyte_B4096.json 2>> $O/hot_b.err
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 16384 --no-interaction > $O/${TAG}_hot_terabyte_B16384.json 2>> $O/hot_b.err
timeout 600 python bench.py > $O/${TAG}_bench_terabyte_n1.json 2> $O/${TAG}_bench_terabyte_n1.err; echo "bench rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_hot_*.json")):
    try:
        r=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split('/')[-1],{k:(round(r[k]['us'],2), round(r[k].get('frac_hbm',0),3)) for k in ('lookup','sort','lookup_sort','sort_plus_update','update_only','embedding_lookup_plus_update','embedding_chain') if k in r})
try:
    r=json.load(open("$O/${TAG}_bench_terabyte_n1.json"))
    print("bench", round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v.get('back_to_back_us',0),2), round(v['in_step_us'],2), round(v.get('frac_hbm',0),3)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r['kernels'].get('_replays'), r.get('embedding'), r.get('e2e_host'), r.get('cpu_baseline'))
except Exception as e:
    print("bench unreadable", e)
PY
tail -5 $O/hot_b.err; tail -5 $O/${TAG}_bench_terabyte_n1.err
