#!/bin/bash
# One GPU-box visit: parity tests, FFMA rate, hot-path microbenchmarks (warp vs tiled interaction), bench, ncu.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r01b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/r01b_pytest_gpu.log
./benchmarks/ffma_rate > $O/r01b_ffma_rate.json 2>&1; cat $O/r01b_ffma_rate.json
python benchmarks/hotpath.py --workload terabyte --B 2048 > $O/r01b_hot_terabyte_B2048.json 2> $O/hot_a.err; cat $O/r01b_hot_terabyte_B2048.json
for k in interaction_fwd interaction_bwd; do
  DLRMB_INTERACT=tiled python benchmarks/hotpath.py --workload terabyte --B 2048 --small-tables --only $k > $O/r01b_hot_terabyte_B2048_tiled_$k.json 2>> $O/hot_a.err; cat $O/r01b_hot_terabyte_B2048_tiled_$k.json
done
python benchmarks/hotpath.py --workload kaggle --B 2048 > $O/r01b_hot_kaggle_B2048.json 2>> $O/hot_a.err; cat $O/r01b_hot_kaggle_B2048.json
python benchmarks/hotpath.py --workload terabyte --B 16384 > $O/r01b_hot_terabyte_B16384.json 2>> $O/hot_a.err; cat $O/r01b_hot_terabyte_B16384.json
python bench.py > $O/r01b_bench_n1.json 2> $O/r01b_bench_n1.err; echo "bench rc=$?"; cat $O/r01b_bench_n1.json; tail -5 $O/r01b_bench_n1.err
for k in interaction_fwd interaction_bwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:interaction -c 2 -f -o $O/r01b_ncu_$k \
    python benchmarks/hotpath.py --workload terabyte --B 2048 --small-tables --only $k --no-graph --iters 1 --nb 2 > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la $O | tail -15
