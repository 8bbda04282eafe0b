#!/bin/bash
# Round 2, GPU visit Q (1 GPU): long-run fast path of the update: parity tests + the skewed / pooled sweep corners + the DLRM case.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r02q}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sgd or update or bf16 or golden or five or terabyte or kaggle" > $O/${TAG}_pytest_update.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_update.log
timeout 300 python benchmarks/hotpath.py --workload terabyte --B 2048 --no-interaction > $O/${TAG}_hot_terabyte_B2048.json 2>> $O/hot_q.err
timeout 300 python benchmarks/hotpath.py --workload kaggle --B 2048 --no-interaction > $O/${TAG}_hot_kaggle_B2048.json 2>> $O/hot_q.err
run() { timeout 300 python benchmarks/hotpath.py --rows $1 --D $2 --B $((1048576 / $3)) --P $3 --zipf $4 --nb 3 --iters 3 --no-interaction > $O/${TAG}_hot_r$1_D$2_P$3_z$4.json 2>> $O/hot_q.err; }
run 10000000 128 16 1.2
run 10000000 128 16 1.05
run 10000000 128 4 1.2
run 10000000 64 16 1.2
run 10000000 64 16 1.05
run 100000 128 16 0.0
run 100000 64 16 1.2
run 10000000 128 1 1.2
run 10000000 128 1 0.0
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_hot_*.json")):
    try:
        r=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split('/')[-1],{k:(round(r[k]['us'],2), round(r[k].get('frac_hbm',0),3)) for k in ('lookup','sort','update_only','embedding_lookup_plus_update','embedding_chain') if k in r})
PY
tail -3 $O/hot_q.err
