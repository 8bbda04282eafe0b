#!/bin/bash
# Round 2, GPU visit P (8 GPUs): A/B at N = 8 -- one-shot peer all-reduce of the bottom MLP gradients vs NCCL, split-dx backward.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02p}
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29530 + RANDOM % 200)) bench.py --gpus 8 --steps 30 --warmup 5 "$@" > $O/${TAG}_bench_n8_$name.json 2> $O/${TAG}_bench_n8_$name.err; echo "bench n8 $name rc=$?"; }
run peer_ar
run nccl_ar --nccl-bottom-allreduce
run peer_ar_split_dx --split-dx
run peer_ar_2
python - <<PY
import json
for f in ("peer_ar","nccl_ar","peer_ar_split_dx","peer_ar_2"):
    try:
        r=json.load(open("$O/${TAG}_bench_n8_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), r['config'].get('exchange_check',{}).get('ok'), r['config'].get('barrier_timeouts'), r['hot_path'].get('timeline_us'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -2 $O/${TAG}_bench_n8_peer_ar.err
