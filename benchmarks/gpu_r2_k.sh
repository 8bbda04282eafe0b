#!/bin/bash
# Round 2, GPU visit K (8 GPUs): bench at N = 8 and N = 4 with the step timeline (tile choice fixed for the owners' updates).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02k}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err; echo "bench n8 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 > $O/${TAG}_bench_n4.json 2> $O/${TAG}_bench_n4.err; echo "bench n4 rc=$?"
python - <<PY
import json
for f in ("bench_n8","bench_n4"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v['in_step_us'],2), round(v.get('event_pair_us',0),2)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r['hot_path'].get('timeline_us'), r['config'].get('barrier_timeouts'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -2 $O/${TAG}_bench_n8.err
