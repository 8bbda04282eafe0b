"""Register-block (TB) / k-split (KS) sweep of the general tiled interaction forward (csrc/interact.cu).
The Criteo shapes default to the warp-per-sample kernels, so the sweep switches the library to the general
kernels through its options (`hotpath.py --opt interact_general=1 --opt fwd_tb=.. --opt fwd_ks=..`)."""
import json
import subprocess
import sys

for wl in ("terabyte", "kaggle"):
    for tb in (3, 6, 9):
        for ks in (0, 1, 2, 3):
            out = subprocess.run([sys.executable, "benchmarks/hotpath.py", "--workload", wl, "--nb", "4", "--iters", "5",
                                  "--only", "interaction_fwd", "--small-tables", "--opt", "interact_general=1",
                                  "--opt", f"fwd_tb={tb}", "--opt", f"fwd_ks={ks}"], capture_output=True, text=True)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print(wl, tb, ks, round(d["interaction_fwd"]["us"], 2), flush=True)
            except Exception:
                print(wl, tb, ks, "ERR", out.stderr[-300:], flush=True)
