#!/bin/bash
# Round 2, GPU visit M (8 GPUs): step map (kernels + MLP blocks + NCCL) at N = 8.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02m}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 8 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err; echo "bench n8 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
for f in ("bench_n8","bench_n1"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]))
        print("  timeline", r['hot_path'].get('timeline_us'))
        m=r['hot_path'].get('step_map_us') or {}
        for k,v in sorted(((k,v) for k,v in m.items() if k!='note'), key=lambda kv: kv[1][0]): print("  ", k, v)
        print("  others", r['hot_path'].get('other_calls_event_pair_us'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -2 $O/${TAG}_bench_n8.err
