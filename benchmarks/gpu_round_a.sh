#!/bin/bash
# One GPU-box visit: parity tests, FFMA / MMA rates, hot-path microbenchmarks (interaction variants), bench, ncu.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01c}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/${TAG}_pytest_gpu.log
./benchmarks/ffma_rate > $O/${TAG}_ffma_rate.json 2>&1; cat $O/${TAG}_ffma_rate.json
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 > $O/${TAG}_hot_terabyte_B2048.json 2> $O/hot_a.err; cat $O/${TAG}_hot_terabyte_B2048.json
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 > $O/${TAG}_hot_kaggle_B2048.json 2>> $O/hot_a.err; cat $O/${TAG}_hot_kaggle_B2048.json
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 16384 > $O/${TAG}_hot_terabyte_B16384.json 2>> $O/hot_a.err; cat $O/${TAG}_hot_terabyte_B16384.json
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"; cat $O/${TAG}_bench_n1.json; tail -5 $O/${TAG}_bench_n1.err
for k in interaction_fwd interaction_bwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:interaction -c 2 -f -o $O/${TAG}_ncu_$k \
    python benchmarks/hotpath.py --workload terabyte --B 2048 --small-tables --only $k --no-graph --iters 1 --nb 2 > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
cat $O/hot_a.err | tail -5
DLRMB_UPDATE_TWO_LAUNCHES=1 timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 --no-interaction > $O/${TAG}_hot_terabyte_B2048_two_launch_update.json 2>> $O/hot_a.err; cat $O/${TAG}_hot_terabyte_B2048_two_launch_update.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"update|lookup_gather|sort_small" -c 9 -f -o $O/${TAG}_ncu_embedding \
    python benchmarks/hotpath.py --workload terabyte --B 2048 --no-graph --iters 1 --nb 2 > $O/ncu_emb.log 2>&1; echo "ncu emb rc=$?"
