"""Multi-GPU checks (skipped on a one-GPU box): BASELINE config 4 driven through the C ABI alone -- NCCL
collectives of dlrmb_comm_*, or the fused peer-store exchanges -- against the unsharded oracle."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")]


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_sharded_step_through_the_c_abi_alone(mode):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cabi_sharded_step.py"), "--gpus", "2", "--mode", mode,
                          "--B", "128", "--steps", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["ok"] and res["pooled_rows_bit_exact"]
    assert res["interaction_fwd_rel_err"] < 1e-5 and res["dx_rel_err"] < 1e-5 and res["tables_rel_err"] < 1e-4
