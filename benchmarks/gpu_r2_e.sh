#!/bin/bash
# Round 2, GPU visit E (1 GPU): regression suite after the p2p fused sort / 1024-thread sort variants,
# hot-path numbers, and an ncu --set full capture of the pooled + skewed update case of the sweep.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02e}
timeout 1800 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 8192 --no-interaction > $O/${TAG}_hot_terabyte_B8192.json 2>> $O/hot_e.err
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 16384 --no-interaction > $O/${TAG}_hot_terabyte_B16384.json 2>> $O/hot_e.err
CMD="python benchmarks/hotpath.py --rows 10000000 --D 128 --B 65536 --P 16 --zipf 1.2 --nb 2 --iters 1 --no-graph --no-interaction"
timeout 300 $CMD > $O/${TAG}_hot_pool16_zipf12.json 2>> $O/hot_e.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"update_tiles|update_fixup|lookup_pool|radix" --launch-skip 20 -c 12 -f -o $O/${TAG}_ncu_pool16 $CMD > $O/ncu_pool.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_hot_*.json")):
    try:
        r=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split('/')[-1],{k:(round(r[k]['us'],2), round(r[k].get('frac_hbm',0),3)) for k in ('lookup','sort','lookup_sort','sort_plus_update','update_only','embedding_lookup_plus_update','embedding_chain') if k in r})
PY
tail -3 $O/hot_e.err
