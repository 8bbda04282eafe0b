"""Register-block (TB) / k-split (KS) sweep of the general tiled interaction forward (csrc/interact.cu).
The Criteo shapes default to the warp-per-sample kernels, so the sweep forces DLRMB_INTERACT=tiled."""
import os, sys, json, subprocess
res = {}
for wl in ("terabyte", "kaggle"):
    for tb in (3, 6, 9):
        for ks in (0, 1, 2, 3):
            env = dict(os.environ, DLRMB_INTERACT="tiled", DLRMB_FWD_TB=str(tb), DLRMB_FWD_KS=str(ks))
            out = subprocess.run([sys.executable, "benchmarks/hotpath.py", "--workload", wl, "--nb", "4", "--iters", "5", "--only", "interaction_fwd", "--small-tables"],
                                 env=env, capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print(wl, tb, ks, round(d["interaction_fwd"]["us"], 2), flush=True)
            except Exception as e:
                print(wl, tb, ks, "ERR", out.stderr[-300:], flush=True)
