#!/bin/bash
# Round 2, GPU visit C (2 GPUs): the sharded path -- C-ABI-only step vs oracle (NCCL and peer-store modes),
# bench at N = 2 with the flag barrier and with the NCCL barrier, N = 1 for the new in-step clock.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02c}
timeout 300 python tests/cabi_sharded_step.py --gpus 2 --mode nccl > $O/${TAG}_cabi_nccl.json 2> $O/${TAG}_cabi_nccl.err; echo "cabi nccl rc=$?"; cat $O/${TAG}_cabi_nccl.json; tail -3 $O/${TAG}_cabi_nccl.err
timeout 300 python tests/cabi_sharded_step.py --gpus 2 --mode p2p > $O/${TAG}_cabi_p2p.json 2> $O/${TAG}_cabi_p2p.err; echo "cabi p2p rc=$?"; cat $O/${TAG}_cabi_p2p.json; tail -3 $O/${TAG}_cabi_p2p.err
timeout 600 python bench.py --no-cpu-baseline --no-host-leg > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
for b in flags nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --barrier $b > $O/${TAG}_bench_n2_$b.json 2> $O/${TAG}_bench_n2_$b.err; echo "bench n2 $b rc=$?"
done
python - <<PY
import json
for f in ("bench_n1","bench_n2_flags","bench_n2_nccl"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v.get('back_to_back_us',0),2), round(v['in_step_us'],2), round(v.get('in_graph_us',0),2)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r.get('embedding',{}).get('frac_hbm'), r['config'].get('exchange_check'), r['config'].get('barrier_timeouts'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -4 $O/${TAG}_bench_n2_flags.err
