#!/bin/bash
# Round 2, final multi-GPU visit: bench at N = 2, 4, 8 (default arguments) and the 2-GPU tests.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02z}
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -1 $O/${TAG}_pytest_multi.log
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n > $O/${TAG}_bench_terabyte_n$n.json 2> $O/${TAG}_bench_terabyte_n$n.err; echo "bench n$n rc=$?"
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --impl reference --steps 10 --warmup 3 > $O/${TAG}_bench_reference_n8.json 2> $O/${TAG}_bench_reference_n8.err; echo "bench ref n8 rc=$?"
python - <<PY
import json
for f in ("bench_terabyte_n2","bench_terabyte_n4","bench_terabyte_n8","bench_reference_n8"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), (r.get('config') or {}).get('exchange_check',{}) and r['config']['exchange_check'].get('ok'), (r.get('config') or {}).get('barrier_timeouts'), r.get('cpu_baseline',{}).get('cores'))
    except Exception as e:
        print(f, "unreadable", e)
PY
