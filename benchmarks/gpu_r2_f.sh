#!/bin/bash
# Round 2, GPU visit F (2 GPUs): bench with the device-clock in-step timing at N = 1 and N = 2 (p2p fused sort,
# early all-reduce of the top MLP's gradients), quick regression of the p2p tests.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02f}
timeout 900 python -m pytest tests/test_gpu_p2p.py -m gpu -q -x > $O/${TAG}_pytest_p2p.log 2>&1; echo "pytest p2p rc=$?"; tail -3 $O/${TAG}_pytest_p2p.log
timeout 600 python bench.py --no-host-leg > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --late-allreduce > $O/${TAG}_bench_n2_late.json 2> $O/${TAG}_bench_n2_late.err; echo "bench n2 late rc=$?"
python - <<PY
import json
for f in ("bench_n1","bench_n2","bench_n2_late"):
    try:
        r=json.load(open("$O/${TAG}_%s.json"%f))
        print(f, round(r["value"]), round(r["ms_per_step"],4), "e2e", round(r["e2e"]["value"]), {k:(round(v.get('back_to_back_us',0),2), round(v['in_step_us'],2), round(v.get('event_pair_us',0),2), round(v.get('frac_hbm',0),3)) for k,v in r.get('kernels',{}).items() if not k.startswith('_')}, r.get('embedding'), r.get('roofline',{}).get('kernel'), r.get('roofline',{}).get('frac'), r.get('cpu_baseline',{}).get('value'))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -4 $O/${TAG}_bench_n2.err
