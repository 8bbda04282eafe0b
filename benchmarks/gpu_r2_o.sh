#!/bin/bash
# Round 2, GPU visit O (1 GPU): is the one-GPU multi-process suite flaky?  Five runs of tests/test_gpu_p2p.py.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02o}
for i in 1 2 3 4 5; do
  timeout 400 python -m pytest tests/test_gpu_p2p.py -m gpu -q -x > $O/${TAG}_pytest_p2p_$i.log 2>&1; echo "run $i rc=$?"; tail -1 $O/${TAG}_pytest_p2p_$i.log
done
