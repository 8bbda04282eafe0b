#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01y}
timeout 300 python benchmarks/fwd_accuracy.py > $O/${TAG}_fwd_accuracy.json 2> $O/acc.err; grep mma $O/${TAG}_fwd_accuracy.json; tail -2 $O/acc.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "interaction or golden or known" > $O/${TAG}_pytest_inter.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_inter.log
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 --small-tables --only interaction_fwd > $O/${TAG}_fwd_B2048.json 2> $O/hot_e.err; python -c "import json; r=json.load(open('$O/${TAG}_fwd_B2048.json')); print('fwd B2048', r['interaction_fwd'])"
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 16384 --small-tables --only interaction_fwd > $O/${TAG}_fwd_B16384.json 2>> $O/hot_e.err; python -c "import json; r=json.load(open('$O/${TAG}_fwd_B16384.json')); print('fwd B16384', r['interaction_fwd'])"
