#!/bin/bash
# Round 2, GPU visit L (2 GPUs): sparse update launched from inside the backward pass: tests, N = 1 and N = 2 bench.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02l}
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
for f in ("bench_n1","bench_n2"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v['in_step_us'],2), round(v.get('frac_hbm',0),3)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r['hot_path'].get('timeline_us'), r.get('embedding',{}).get('frac_hbm'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $O/${TAG}_bench_n2.err
