"""Fused lookup + peer-store exchange (csrc/p2p.cu) with two processes sharing cuda:0.

CUDA IPC works between processes on one device, so the kernel, the slot map and the IPC mapping
are exercised without a second GPU; the cross-rank ordering here is a device synchronize + a gloo
barrier (NCCL refuses two ranks on one GPU).  The multi-GPU wiring itself (NCCL barrier, autograd)
is checked by bench.py at start-up on every multi-GPU run (`exchange_check` in its JSON line).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

ROWS = [50, 7, 400, 3, 1200, 33, 9]
D, BL, P = 64, 48, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        from dlrm_jl_b200.sharded import ShardedEmbedding
        from oracle import oracle as O
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        se = ShardedEmbedding.create(ROWS, D, BL, P, rank, world, dev)

        def barrier():
            torch.cuda.synchronize()
            dist.barrier()

        se.enable_peer_exchange(BL, barrier=barrier)
        mine = se.local_ids
        local_np = {k: se.tables.download(j) for j, k in enumerate(mine)}
        gathered = [None] * world
        dist.all_gather_object(gathered, local_np)
        tables = [None] * len(ROWS)
        for part in gathered:
            for k, v in part.items():
                tables[k] = v
        rng = np.random.default_rng(5)
        idx_all = [np.stack([rng.integers(0, r, size=(BL, P)) for r in ROWS]) for _ in range(world)]
        # owner view: my tables, every rank's samples (rank-major)
        idx_owned = np.stack([np.concatenate([idx_all[r][k] for r in range(world)], axis=0) for k in mine])
        se.tables.lookup_p2p(torch.from_numpy(idx_owned).to(dev), se.peer.peer_ptrs, BL, 1 + len(ROWS))
        se.peer.barrier()
        T = se.peer.T.cpu().numpy()
        ref = O.lookup(tables, list(idx_all[rank]), slot0=1)
        ok = bool(np.array_equal(T[:, 1:], ref[:, 1:]))
        barrier()
        q.put((rank, ok, ""))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, False, repr(exc)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_fused_lookup_peer_store_two_ranks_one_gpu():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in res:
        assert ok, f"rank {rank}: pooled rows stored by peers differ from the oracle lookup {err}"
