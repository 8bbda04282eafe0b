"""World-size-2/3 CPU tests (gloo) of the table-sharded plumbing: partitioning, the three
all-to-alls and the dense all-reduce.  The local compute is injected; here it is the CPU oracle
(allowed in tests only), so the result can be compared with the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlrm_jl_b200.sharded import ShardedEmbedding, TableSharding, allreduce_dense_grads

ROWS = [50, 7, 400, 3, 120, 33, 9]
D, BL, P = 8, 6, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_partition_balances_lookups_then_bytes():
    from dlrm_jl_b200.model import TERABYTE_EMBEDDING_SIZES
    rows = [min(r, 40_000_000) for r in TERABYTE_EMBEDDING_SIZES]
    sh = TableSharding.build(rows, 8)
    counts = sh.counts()
    assert sorted(counts) == [3, 3, 3, 3, 3, 3, 4, 4]
    big = [k for k, r in enumerate(rows) if r == 40_000_000]
    assert len({sh.owner[k] for k in big}) == len(big) == 5      # the 20.5 GB tables on different GPUs
    assert max(sh.bytes_per_rank(128)) < 25e9
    assert sorted(k for l in sh.local for k in l) == list(range(26))
    for w in (1, 2, 4):
        c = TableSharding.build(rows, w).counts()
        assert max(c) - min(c) <= 1


def test_c_abi_shard_plan_equals_the_python_plan():
    """dlrmb_shard_plan (host logic of libdlrm_b200.so, what a non-Python host calls) deals the tables exactly as
    TableSharding.build does -- descending rows, ties by table id, snake order -- for the Criteo geometries, for
    ties and for more ranks than tables; bad arguments are refused."""
    import ctypes as C
    from dlrm_jl_b200 import _lib
    from dlrm_jl_b200.model import KAGGLE_EMBEDDING_SIZES, TERABYTE_EMBEDDING_SIZES
    lib = _lib.load()
    rng = np.random.default_rng(3)
    cases = [list(KAGGLE_EMBEDDING_SIZES), [min(r, 40_000_000) for r in TERABYTE_EMBEDDING_SIZES], [7, 7, 7, 7, 7],
             [1], [5, 1000, 5, 1000, 3]] + [list(rng.integers(1, 10 ** 7, size=int(n))) for n in rng.integers(1, 40, size=6)]
    for rows in cases:
        for world in (1, 2, 3, 4, 8, 16):
            owner = (C.c_int32 * len(rows))()
            _lib.check(lib.dlrmb_shard_plan(len(rows), (C.c_int64 * len(rows))(*[int(r) for r in rows]), world, owner))
            assert list(owner) == TableSharding.build(rows, world).owner, (rows, world)
    assert lib.dlrmb_shard_plan(0, None, 2, None) == _lib.EINVAL
    assert lib.dlrmb_shard_plan(3, (C.c_int64 * 3)(1, 2, 3), 0, (C.c_int32 * 3)()) == _lib.EINVAL


def _worker(rank, world, port, q):
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)                      # same tables / batches on every rank
        tables = [rng.standard_normal((r, D)).astype(np.float32) for r in ROWS]
        idx_all = [np.stack([rng.integers(0, r, size=(BL, P)) for r in ROWS]) for _ in range(world)]
        x_all = [rng.standard_normal((BL, D)).astype(np.float32) for _ in range(world)]
        gz_all = [rng.standard_normal((BL, D + 28)).astype(np.float32) for _ in range(world)]
        sh = TableSharding.build(ROWS, world)
        mine = sh.local[rank]
        local_tables = [tables[k].copy() for k in mine]

        def lookup_fn(idx, out, slot0):
            out.copy_(torch.from_numpy(O.lookup(local_tables, list(idx.numpy()), slot0=slot0)))

        def update_fn(idx, g, lr, slot0, presorted=False):
            for j in range(len(mine)):
                O.sparse_sgd_update_fast(local_tables[j], idx[j].numpy(), np.ascontiguousarray(g[:, slot0 + j].numpy()), lr)

        se = ShardedEmbedding(ROWS, D, rank, world, lookup_fn, update_fn)
        anchor = torch.zeros(1, requires_grad=True)
        T = se.lookup(torch.from_numpy(idx_all[rank]), anchor)
        # forward parity: every rank sees the pooled rows of ITS samples for ALL tables
        ref_T = O.lookup(tables, list(idx_all[rank]), slot0=1)
        ok_fwd = np.array_equal(T.detach().numpy()[:, 1:], ref_T[:, 1:])
        # backward: interaction pullback on the oracle, routed back through autograd
        Tn = T.detach().numpy().copy()
        Tn[:, 0] = x_all[rank]
        _dx, dT = O.interaction_bwd(gz_all[rank], Tn)
        T.backward(torch.from_numpy(dT))
        se.update(0.5)
        # dense all-reduce
        p = torch.nn.Parameter(torch.zeros(5))
        p.grad = torch.full((5,), float(rank + 1))
        allreduce_dense_grads([p], world)
        ok_ar = bool(torch.allclose(p.grad, torch.full((5,), sum(range(1, world + 1)) / world)))
        # unsharded reference: one process applying every rank's updates to the same tables
        ref = [t.copy() for t in tables]
        idx_glob = [np.concatenate([idx_all[r][k] for r in range(world)], axis=0) for k in range(len(ROWS))]
        dT_glob = []
        for r in range(world):
            Tr = O.lookup(tables, list(idx_all[r]), slot0=1)
            Tr[:, 0] = x_all[r]
            dT_glob.append(O.interaction_bwd(gz_all[r], Tr)[1])
        dT_glob = np.concatenate(dT_glob, axis=0)
        for k in range(len(ROWS)):
            O.sparse_sgd_update_fast(ref[k], idx_glob[k], np.ascontiguousarray(dT_glob[:, 1 + k]), 0.5)
        ok_upd = all(np.array_equal(local_tables[j], ref[k]) for j, k in enumerate(mine))
        q.put((rank, ok_fwd, ok_upd, ok_ar))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_step_matches_unsharded_oracle(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_fwd, ok_upd, ok_ar in res:
        assert ok_fwd, f"rank {rank}: forward exchange mismatch"
        assert ok_upd, f"rank {rank}: sharded update differs from unsharded oracle"
        assert ok_ar, f"rank {rank}: dense all-reduce mismatch"
