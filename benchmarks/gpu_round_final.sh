#!/bin/bash
# Final single-GPU measurements for the round: smoke, tests, microbenchmarks, bench (both workloads), ncu.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01z}
timeout 300 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_gpu.log
./benchmarks/ffma_rate > $O/${TAG}_ffma_mma_rate.json 2>&1
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 > $O/${TAG}_hot_terabyte_B2048.json 2> $O/hot_z.err
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 16384 > $O/${TAG}_hot_terabyte_B16384.json 2>> $O/hot_z.err
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 > $O/${TAG}_hot_kaggle_B2048.json 2>> $O/hot_z.err
timeout 600 python bench.py > $O/${TAG}_bench_terabyte_n1.json 2> $O/${TAG}_bench_terabyte_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload kaggle --no-cpu-baseline > $O/${TAG}_bench_kaggle_n1.json 2> $O/${TAG}_bench_kaggle_n1.err; echo "bench kaggle rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/${TAG}_bench_reference_cpu.json 2> $O/${TAG}_bench_reference_cpu.err; echo "bench ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"interaction|update_tiles|lookup_gather|sort_small|dense_bwd|bce" --launch-skip 40 -c 12 -f -o $O/${TAG}_ncu_step \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 26 -c 1200 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import json
for f in ("hot_terabyte_B2048","hot_terabyte_B16384","hot_kaggle_B2048"):
    r=json.load(open("$O/${TAG}_%s.json"%f))
    print(f,{k:(round(r[k]['us'],2), round(r[k].get('frac_hbm',0),3)) for k in ('lookup','sort','sort_plus_update','update_only','embedding_lookup_plus_update','interaction_fwd','interaction_bwd') if k in r})
for f in ("bench_terabyte_n1","bench_kaggle_n1","bench_reference_cpu"):
    r=json.load(open("$O/${TAG}_%s.json"%f))
    print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v['avg_us'],2), round(v.get('frac_hbm',0),3)) for k,v in r.get('kernels',{}).items()}, r.get('roofline',{}).get('kernel'), r.get('roofline',{}).get('frac'), r.get("cpu_baseline",{}).get("value"))
PY
