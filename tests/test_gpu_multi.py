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


def _ar_worker(rank, world, port, q):
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(200, exit=True)      # never hang the suite
    try:
        import numpy as np
        import torch.distributed as dist
        from dlrm_jl_b200.sharded import PeerExchange
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        peer = PeerExchange(8, 3, 16, rank, world, dev)          # small T buffer; flag barrier on
        n = 4 * 1000
        peer.enable_small_allreduce(n)
        ok = True
        for step in range(5):                                    # several steps: both halves of the buffer, epochs
            rngs = [np.random.default_rng(10 * step + r) for r in range(world)]
            parts = [g.standard_normal(n).astype(np.float32) for g in rngs]
            t = torch.from_numpy(parts[rank]).to(dev)
            peer.allreduce_small(t)
            ref = parts[0].copy()
            for r in range(1, world):
                ref = (ref + parts[r]).astype(np.float32)         # rank order, fp32: the kernel's order
            ok = ok and bool(np.array_equal(t.cpu().numpy(), ref))
            peer.barrier(0)                                       # the flag barrier on its own, too
        ok = ok and peer.barrier_timeouts() == 0
        torch.cuda.synchronize()
        dist.barrier()
        peer.close()
        q.put((rank, ok, ""))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, False, repr(exc) + traceback.format_exc()))
    finally:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def test_peer_allreduce_and_flag_barrier_two_gpus():
    """dlrmb_peer_allreduce_f32 / dlrmb_peer_barrier with one rank per GPU: rank-ordered fp32 sums, identical on
    every rank, over several steps (both halves of the double buffer), no barrier time-outs."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ar_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    res = []
    try:
        for _ in range(world):
            try:
                res.append(q.get(timeout=240))
            except queue.Empty:
                res.append((-1, False, "a worker did not answer within 240 s"))
                break
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.terminate()
    for rank, ok, err in res:
        assert not err, f"rank {rank}: {err}"
        assert ok, f"rank {rank}: peer all-reduce differs from the rank-ordered sum"
