"""
Multi-process host logic on CPU (gloo, world size 2): batch sharding, gathering and the
max-over-ranks timing reduction used by bench.py.  The per-rank compute is a stand-in callable
(the CUDA kernels need a GPU); on a GPU box the same code path runs with NCCL.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exponax_b200 import _distributed as D


def test_shard_bounds_cover_and_balance():
    for batch in (0, 1, 2, 7, 16, 16384, 16385):
        for ws in (1, 2, 3, 4, 8):
            bounds = [D.shard_bounds(batch, ws, r) for r in range(ws)]
            assert bounds[0][0] == 0 and bounds[-1][1] == batch
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        u0 = np.arange(batch * 3, dtype=np.float32).reshape(batch, 1, 3)

        def fake_rollout(u):  # (b, C, N) -> (b, T=2, C, N), per-trajectory independent
            return np.stack([u + 1, u + 2], axis=1)

        local = D.sharded_apply(fake_rollout, u0)
        lo, hi = D.shard_bounds(batch, ws, rank)
        assert local.shape == (hi - lo, 2, 1, 3)
        np.testing.assert_array_equal(local, fake_rollout(u0[lo:hi]))
        full = D.sharded_apply(fake_rollout, u0, gather=True)
        np.testing.assert_array_equal(full, fake_rollout(u0))
        np.testing.assert_array_equal(D.local_shard(u0), u0[lo:hi])
        t = D.max_over_ranks(1.0 + rank)
        assert t == float(ws)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 8])
def test_sharded_apply_gloo_world2(batch):
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, batch, q)) for r in range(ws)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_single_process_passthrough():
    u0 = np.ones((3, 1, 4), np.float32)
    assert D.world() == (0, 1)
    out = D.sharded_apply(lambda u: u * 2, u0, gather=True)
    np.testing.assert_array_equal(out, u0 * 2)
    assert D.max_over_ranks(3.5) == 3.5


# ---------------------------------------------------------------- slab transposes (config c5)
def _slab_worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from exponax_b200._slab import transpose_a_to_b, transpose_b_to_a
        F, N, K = 3, 8, 5
        n = N // ws
        g = torch.arange(F * N * N * K, dtype=torch.float32).reshape(F, N, N, K)
        g = torch.complex(g, -g)
        from exponax_b200._spectral import SLAB_CYCLIC, slab_indices
        a = g[:, rank * n:(rank + 1) * n].contiguous()       # physical slab: split over axis 0 (contiguous x-planes)
        b = g[:, :, slab_indices(rank, ws, N)].contiguous()  # spectral slab: split over axis 1 (cyclic by default)
        owner = (lambda i: (i % ws, i // ws)) if SLAB_CYCLIC else (lambda i: (i // n, i % n))
        assert torch.equal(transpose_a_to_b(a), b)
        assert torch.equal(transpose_b_to_a(b), a)
        assert torch.equal(transpose_b_to_a(transpose_a_to_b(a)), a)
        # raw exchange (EXB_SLAB_SEGMENTED): the all-to-all lands in peer-major order [src][x][k1 within src][K];
        # entry i of an axis-1 line lives at (i // n) * (n*n*K) + x * (n*K) + (i % n) * K  -- what the segmented
        # column pass and the peer-store kernels address
        from exponax_b200._slab import exchange_a_to_b_raw, exchange_b_to_a_raw
        for f in range(F):
            raw = torch.empty_like(a[f])
            exchange_b_to_a_raw(b[f], raw)
            flat = raw.reshape(-1)
            for x in range(n):
                for i in range(N):   # entry i of an axis-1 line: block [owner rank] of the buffer, local position
                    r, loc = owner(i)
                    off = r * (n * n * K) + x * (n * K) + loc * K
                    assert torch.equal(flat[off:off + K], a[f, x, i])
            back = torch.empty_like(b[f])
            exchange_a_to_b_raw(raw, back)
            assert torch.equal(back, b[f])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_slab_transposes_gloo_world2():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slab_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
