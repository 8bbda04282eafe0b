#!/bin/bash
# Round 2, GPU visit N (2 GPUs): one-shot peer all-reduce: 2-GPU tests, N = 2 bench with and without.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02n}
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_p2p.py -m gpu -q -x > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 2 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29526 bench.py --gpus 2 --nccl-bottom-allreduce > $O/${TAG}_bench_n2_nccl.json 2> $O/${TAG}_bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
python - <<PY
import json
for f in ("bench_n2","bench_n2_nccl"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), r['config'].get('exchange_check',{}).get('ok'), r['config'].get('barrier_timeouts'))
        m=r['hot_path'].get('step_map_us') or {}
        for k,v in sorted(((k,v) for k,v in m.items() if k!='note'), key=lambda kv: kv[1][0]): print("  ", k, v)
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $O/${TAG}_bench_n2.err
