#!/bin/bash
# Round 2, GPU visit D (1 GPU): globaltimer resolution, config-5 sweep through bench.py, ncu launch list and
# one --set full capture of the step's kernels.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02d}
./benchmarks/gtimer > $O/${TAG}_globaltimer.json; cat $O/${TAG}_globaltimer.json
timeout 1200 python bench.py --workload sweep > $O/${TAG}_sweep.json 2> $O/${TAG}_sweep.err; echo "sweep rc=$?"; tail -2 $O/${TAG}_sweep.err
CMD="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-host-leg"
timeout 600 $CMD > $O/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 26 -c 1400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 $CMD > $O/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interaction|update_tiles|lookup_sort|lookup_gather|sort_small|bce" --launch-skip 40 -c 10 -f -o $O/${TAG}_ncu_step $CMD > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
python - <<PY
import json
r=json.load(open("$O/${TAG}_sweep.json"))
print(r["summary"], r["value"])
PY
ls -la $O | tail -8
