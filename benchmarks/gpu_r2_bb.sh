#!/bin/bash
# Round 2, GPU visit BB (1 GPU, short): parity of every backward variant, cold A/B of the resident kernel with S stored
# once (FFMA2 with a scalar operand, 3) and the streaming kernels (5, 6) across batch sizes.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02bb}
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "register_variants" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
timeout 120 python benchmarks/ab_bwd.py --variants 2 3 5 6 --B 2048 4096 8192 16384 --iters 6 > $O/${TAG}_ab_bwd.jsonl 2> $O/${TAG}_ab_bwd.err; echo "ab bwd rc=$?"; cut -c1-140 $O/${TAG}_ab_bwd.jsonl
