#!/bin/bash
# Round 2, last visit (1 GPU): ncu launch list and one --set full capture of the final step's kernels.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02z}
CMD="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-host-leg"
timeout 600 $CMD > $O/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 26 -c 1400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 $CMD > $O/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interaction|update_tiles|lookup_sort|bce" --launch-skip 30 -c 10 -f -o $O/${TAG}_ncu_step $CMD > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
ls -la $O | grep ${TAG}_ | tail -6
