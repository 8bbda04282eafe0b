#!/bin/bash
# Round 2, GPU visit AA (1 GPU): one ncu --set full capture of the hot-path kernels launched alone on cold inputs
# (cta_timeline.py, one repetition): lookup+sort, update, forward, backward variants 1 / 2 / 5.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02aa}
CMD="python benchmarks/cta_timeline.py --reps 1 --bwd-variants 1 2 5"
timeout 200 $CMD > $O/${TAG}_plain.log 2>&1; echo "plain rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"interaction_bwd|interaction_fwd|update_tiles|lookup_sort" -c 16 -f -o $O/${TAG}_ncu_kernels $CMD > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O | grep ${TAG}_ | tail -4
