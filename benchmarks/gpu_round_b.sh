#!/bin/bash
# Update-kernel tile sweep at DLRM batch sizes + ncu of the embedding kernels.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
TAG=${TAG:-r01e}
for wl in terabyte kaggle; do
for tile in 4 8 16; do
  DLRMB_UPDATE_TILE=$tile timeout 300 python benchmarks/hotpath.py --workload $wl --B 2048 --no-interaction > $O/${TAG}_hot_${wl}_B2048_tile$tile.json 2>> $O/hot_b.err
  python - <<PY
import json
r=json.load(open("$O/${TAG}_hot_${wl}_B2048_tile$tile.json"))
print("$wl tile $tile", {k: round(r[k]["us"],2) for k in ("lookup","sort","sort_plus_update","update_only")})
PY
done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"update" -c 3 -f -o $O/${TAG}_ncu_update \
    python benchmarks/hotpath.py --workload terabyte --B 2048 --no-graph --iters 1 --nb 2 --no-interaction > $O/ncu_upd.log 2>&1; echo "ncu upd rc=$?"
tail -3 $O/hot_b.err
