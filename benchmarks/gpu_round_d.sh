#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01i}
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/${TAG}_bench_fused.json 2> $O/${TAG}_bench_fused.err; echo "bench fused rc=$?"; tail -3 $O/${TAG}_bench_fused.err
timeout 600 python bench.py --no-cpu-baseline --unfused-mlp > $O/${TAG}_bench_unfused.json 2> $O/${TAG}_bench_unfused.err; echo "bench unfused rc=$?"
python - <<PY
import json
for n in ("fused","unfused"):
    try:
        r=json.load(open("$O/${TAG}_bench_%s.json"%n))
        print(n, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), r["e2e"]["last_loss"], r["gpu_launches"])
    except Exception as e: print(n, "failed", e)
PY
