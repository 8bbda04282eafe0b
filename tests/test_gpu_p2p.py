"""The table-sharded CUDA path (BASELINE config 4) with two or three processes sharing cuda:0.

CUDA IPC works between processes on one device, so the peer-store kernels (csrc/p2p.cu, the scattering
interaction backward), the slot maps and the IPC mappings are exercised without a second GPU; the
cross-rank ordering here is a device synchronize + a gloo barrier on the host (NCCL refuses two ranks
on one GPU, and no kernel of one process ever waits for a kernel of another).  The full sharded
training step -- fused forward exchange, interaction forward / backward with the gradient scatter,
owner-side sort + sparse SGD -- is compared with the UNSHARDED oracle.  On real multi-GPU runs
bench.py repeats the same oracle comparison at start-up (`exchange_check` in its JSON line).
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


def _arm_watchdog(seconds: int = 200) -> None:
    """A worker that is stuck dumps every thread's stack to stderr and exits instead of hanging the suite."""
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(seconds, exit=True)


def _collect(procs, q, world, timeout):
    """Results of the workers; whatever happens, no worker survives the test."""
    import queue
    res = []
    try:
        for _ in range(world):
            try:
                res.append(q.get(timeout=timeout))
            except queue.Empty:
                res.append((-1, False, {}, "a worker did not answer within %d s (see its stack dump on stderr)" % timeout))
                break
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.terminate()
                p.join(timeout=10)
    return res


def _worker(rank, world, port, q):
    _arm_watchdog()
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
    res = _collect(procs, q, world, 240)
    for r in res:
        rank, ok, err = r[0], r[1], r[-1]
        assert ok, f"rank {rank}: pooled rows stored by peers differ from the oracle lookup {err}"


def _step_worker(rank, world, port, q, D, BL, P, rows, steps):
    _arm_watchdog()
    try:
        from dlrm_jl_b200.interact import DotInteraction, interaction_width
        from dlrm_jl_b200.sharded import ShardedEmbedding
        from oracle import oracle as O
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        ntab, F = len(rows), len(rows) + 1
        se = ShardedEmbedding.create(rows, D, BL, P, rank, world, dev)

        def barrier():
            torch.cuda.synchronize()
            dist.barrier()

        # index exchange by peer stores too (int64 ids here); the flag barrier itself needs one GPU per
        # rank, so the ordering is the host-side barrier above
        se.enable_peer_exchange(BL, barrier=barrier, P=P, idx_bytes=8)
        se.enable_fused_backward(BL, split_dx=(D == 128))      # the bench's form: dx first, peer stores on a side stream
        mine = se.local_ids

        def gather_tables():
            gathered = [None] * world
            dist.all_gather_object(gathered, {k: se.tables.download(j) for j, k in enumerate(mine)})
            out = [None] * ntab
            for part in gathered:
                for k, v in part.items():
                    out[k] = v
            return out

        dot = DotInteraction()
        w = interaction_width(F, D)
        worst = {"T": 0.0, "z": 0.0, "dx": 0.0, "tables": 0.0}
        ok_T = True
        lr = 0.25
        for step in range(steps):
            # every step starts from the GPU's own tables (the update is checked to 1e-4, not bit for bit,
            # so the oracle's copy must not drift away from what the next lookup reads)
            ref_tables = gather_tables()
            rng = np.random.default_rng(100 + step)            # same stream on every rank
            idx_all = [np.stack([rng.integers(0, r, size=(BL, P)) for r in rows]) for _ in range(world)]
            x_all = [rng.standard_normal((BL, D)).astype(np.float32) for _ in range(world)]
            gz_all = [rng.standard_normal((BL, w)).astype(np.float32) for _ in range(world)]
            # ---- sharded CUDA step
            x = torch.from_numpy(x_all[rank]).to(dev).requires_grad_(True)
            T = se.lookup_fused(torch.from_numpy(idx_all[rank]).to(dev))
            se.sort_async()
            z = dot(x, T, scatter=se.scatter_plan)
            if D == 16:     # the bench's form: barrier + update launched from inside the backward pass, on a side stream
                side = torch.cuda.Stream()
                se.update_inside_backward(lr, side)
                z.backward(torch.from_numpy(gz_all[rank]).to(dev))
                torch.cuda.current_stream().wait_stream(side)
            else:
                z.backward(torch.from_numpy(gz_all[rank]).to(dev))
                se.finish_backward()
                se.update(lr)
            barrier()
            # ---- unsharded oracle step on the same inputs
            T_ref = O.lookup(ref_tables, list(idx_all[rank]), slot0=1)
            ok_T = ok_T and bool(np.array_equal(T.cpu().numpy()[:, 1:], T_ref[:, 1:]))
            T_ref[:, 0] = x_all[rank]
            ok_T = ok_T and bool(np.array_equal(T.cpu().numpy()[:, 0], x_all[rank]))     # fused fast_vcat
            worst["z"] = max(worst["z"], O.rel_err(z.detach().cpu().numpy(), O.interaction_fwd(T_ref)))
            dT_glob = []
            for r in range(world):
                Tr = O.lookup(ref_tables, list(idx_all[r]), slot0=1)
                Tr[:, 0] = x_all[r]
                dx_r, dT_r = O.interaction_bwd(gz_all[r], Tr)
                dT_glob.append(dT_r)
                if r == rank:
                    worst["dx"] = max(worst["dx"], O.rel_err(x.grad.cpu().numpy(), dx_r))
            dT_glob = np.concatenate(dT_glob, axis=0)
            for k in range(ntab):
                idx_glob = np.concatenate([idx_all[r][k] for r in range(world)], axis=0)
                O.sparse_sgd_update_fast(ref_tables[k], idx_glob, np.ascontiguousarray(dT_glob[:, 1 + k]), lr)
            for j, k in enumerate(mine):
                worst["tables"] = max(worst["tables"], O.rel_err(se.tables.download(j), ref_tables[k]))
            barrier()
        barrier()
        se.close()
        q.put((rank, ok_T, worst, ""))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, False, {}, repr(exc) + traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world,D,BL,P,rows", [
    (2, 128, 96, 1, [50, 7, 400, 3, 1200, 33, 9] + [20 + 31 * k for k in range(19)]),   # F = 27, d = 128: the bench kernels
    (2, 64, 48, 2, [50, 7, 400, 3, 1200, 33, 9]),                                         # F = 8: general tiled kernels, pooling
    (3, 16, 40, 1, [50, 7, 400, 3, 1200, 33, 9]),                                         # F = 8, d = 16 specialisation, 3 ranks
])
def test_sharded_training_step_matches_unsharded_oracle(world, D, BL, P, rows):
    """Full sharded step on CUDA (config 4's path) vs the unsharded oracle: pooled rows bit-exact,
    interaction output / dx rel 1e-5, every owner's tables rel 1e-4 after two steps."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, q, D, BL, P, rows, 2)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world, 240)
    for rank, ok_T, worst, err in res:
        assert not err, f"rank {rank}: {err}"
        assert ok_T, f"rank {rank}: pooled rows / fused fast_vcat differ from the oracle"
        assert worst["z"] < 1e-5 and worst["dx"] < 1e-5, (rank, worst)
        assert worst["tables"] < 1e-4, (rank, worst)
