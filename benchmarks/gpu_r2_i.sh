#!/bin/bash
# Round 2, GPU visit I (2 GPUs): split-dx backward (tests + N = 2 bench with and without), regression of the suite.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02i}
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --no-split-dx > $O/${TAG}_bench_n2_nosplit.json 2> $O/${TAG}_bench_n2_nosplit.err; echo "bench n2 nosplit rc=$?"
python - <<PY
import json
for f in ("bench_n2","bench_n2_nosplit"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v['in_step_us'],2), round(v.get('event_pair_us',0),2)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r['config'].get('exchange_check',{}).get('ok'), r['config'].get('barrier_timeouts'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -4 $O/${TAG}_bench_n2.err
