#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01u}
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench_n1.err
python - <<PY
import json
r=json.load(open("$O/${TAG}_bench_n1.json"))
print(round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), r["e2e"]["last_loss"], r["gpu_launches"])
PY
