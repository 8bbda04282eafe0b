#!/bin/bash
# Round 2, GPU visit G (1 GPU): tcgen05.mma issue-rate probe, packed-S backward (correctness + speed).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02g}
timeout 120 ./benchmarks/umma_rate > $O/${TAG}_umma_rate.json 2> $O/${TAG}_umma_rate.err; echo "umma rc=$?"; cat $O/${TAG}_umma_rate.json; cat $O/${TAG}_umma_rate.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "packed or backward or ksplit or forward" > $O/${TAG}_pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_bwd.log
for B in 2048 16384; do
  timeout 300 python benchmarks/hotpath.py --workload terabyte --B $B --small-tables --only interaction_bwd --nb 16 > $O/${TAG}_bwd_dup_B$B.json 2>> $O/hot_g.err
  timeout 300 python benchmarks/hotpath.py --workload terabyte --B $B --small-tables --only interaction_bwd --nb 16 --opt bwd_packed=1 > $O/${TAG}_bwd_packed_B$B.json 2>> $O/hot_g.err
done
for B in 2048 16384; do
  timeout 300 python benchmarks/hotpath.py --workload terabyte --B $B --small-tables --only interaction_fwd --nb 16 > $O/${TAG}_fwd_base_B$B.json 2>> $O/hot_g.err
  timeout 300 python benchmarks/hotpath.py --workload terabyte --B $B --small-tables --only interaction_fwd --nb 16 --opt fwd_ksplit=1 > $O/${TAG}_fwd_ksplit_B$B.json 2>> $O/hot_g.err
done
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 --small-tables --only interaction_fwd --nb 16 > $O/${TAG}_fwd_base_kaggle.json 2>> $O/hot_g.err
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 --small-tables --only interaction_fwd --nb 16 --opt fwd_ksplit=1 > $O/${TAG}_fwd_ksplit_kaggle.json 2>> $O/hot_g.err
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 --small-tables --only interaction_bwd --nb 16 > $O/${TAG}_bwd_dup_kaggle.json 2>> $O/hot_g.err
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 --small-tables --only interaction_bwd --nb 16 --opt bwd_packed=1 > $O/${TAG}_bwd_packed_kaggle.json 2>> $O/hot_g.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_bwd_*.json")) + sorted(glob.glob("$O/${TAG}_fwd_*.json")):
    try:
        r=json.load(open(f)); k='interaction_bwd' if 'interaction_bwd' in r else 'interaction_fwd'; print(f.split('/')[-1], round(r[k]['us'],2), round(r[k]['frac_hbm'],3))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $O/hot_g.err
