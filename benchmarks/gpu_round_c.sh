#!/bin/bash
# Launch list of the bench step (ncu, one metric) + interaction microbenchmarks.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01h}
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 --small-tables --only interaction_fwd > $O/${TAG}_fwd_B2048.json 2> $O/hot_c.err; cat $O/${TAG}_fwd_B2048.json
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 16384 --small-tables --only interaction_fwd > $O/${TAG}_fwd_B16384.json 2>> $O/hot_c.err; cat $O/${TAG}_fwd_B16384.json
timeout 600 python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > $O/${TAG}_bench_nograph.json 2> $O/${TAG}_bench_nograph.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 26 -c 2500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
wc -l $O/${TAG}_launches.csv
