"""World-size-2 (and 3) checks of the multi-GPU partitioning on CPU tensors over gloo: partition arithmetic, halo
exchange, r x r all-reduce and the result gather.  The compute between the communication steps is done with numpy
here (tests only); on the GPU box the same steps run between the CUDA kernels (morfem_b200.dist.ShardedHotPath)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from morfem_b200 import dist as mfd
from morfem_b200 import synthetic


def test_even_split_covers_everything():
    for total in (0, 1, 7, 10000, 200001):
        for world in (1, 2, 3, 8):
            r = mfd.owner_ranges(total, world)
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mfd.even_split(10, 2, 2)


def test_halo_plan_is_symmetric():
    ct, _ = synthetic.waveguide_operators(4, 3, 50)
    n = ct.shape[0]
    for world in (2, 3, 4):
        windows = [mfd.column_window(ct.indptr, ct.indices, *mfd.even_split(n, world, p), n) for p in range(world)]
        plans = [mfd.build_halo_plan(p, world, n, windows) for p in range(world)]
        half_bw = 4 * 3 + 4 + 1
        for p, plan in enumerate(plans):
            lo, hi = mfd.even_split(n, world, p)
            assert max(0, lo - half_bw) <= plan.win0 <= lo and hi <= plan.win1 <= min(n, hi + half_bw)
            assert plan.win0 < lo or p == 0
            for peer, glo, ghi in plan.recv:
                assert (p, glo, ghi) in plans[peer].send      # every receive has its matching send
            assert plan.halo_rows == (lo - plan.win0) + (plan.win1 - hi)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_grid, r, f_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ct, tt = synthetic.waveguide_operators(*n_grid)
        n = ct.shape[0]
        rng = np.random.default_rng(0)
        s = rng.standard_normal((n, r)) + 1j * rng.standard_normal((n, r))
        row0, row1 = mfd.even_split(n, world, rank)
        s_loc = torch.from_numpy(s[row0:row1].copy())
        # stage 1: partial Gram + all-reduce
        g = s_loc.conj().T @ s_loc
        mfd.allreduce_sum_(g)
        assert np.allclose(g.numpy(), s.conj().T @ s, rtol=1e-12, atol=1e-9)
        # stage 2: halo exchange, local SpMM on the window, partial projection + all-reduce
        window = mfd.column_window(ct.indptr, ct.indices, row0, row1, n)
        plan = mfd.build_halo_plan(rank, world, n, mfd.gather_windows(window))
        win = mfd.exchange_halo(s_loc, plan)
        assert np.array_equal(win.numpy(), s[plan.win0:plan.win1])
        # the same window through the slab all-gather (the exchange a CUDA graph can capture)
        windows = mfd.gather_windows(window)
        owned = mfd.owner_ranges(n, world)
        h = mfd.slab_halo_rows(windows, owned)
        assert h is not None and 0 < h <= 3 * 2 + 3 + 1
        win2 = mfd.exchange_halo_slabs(s_loc, plan, h, owned)
        assert np.array_equal(win2.numpy(), s[plan.win0:plan.win1])
        at = ct.T.tocsr()
        y_loc = at[row0:row1, plan.win0:plan.win1] @ win.numpy()
        a_r = torch.from_numpy(y_loc.T @ s[row0:row1])
        mfd.allreduce_sum_(a_r)
        assert np.allclose(a_r.numpy(), (s.T @ ct) @ s, rtol=1e-12, atol=1e-9)
        # stage 3/4: contiguous point blocks, gather restores the global order
        f0, f1 = mfd.even_split(f_total, world, rank)
        local = torch.arange(f0, f1, dtype=torch.float64)[:, None, None] * torch.ones((1, 2, 2), dtype=torch.complex128) * (1 + 2j)
        allp = mfd.gather_points(local, f_total)
        expect = np.arange(f_total)[:, None, None] * np.ones((1, 2, 2)) * (1 + 2j)
        assert allp.shape == (f_total, 2, 2) and np.array_equal(allp.numpy(), expect)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,f_total", [(2, 11), (2, 10), (3, 10)])
def test_sharded_steps_over_gloo(tmp_path, world, f_total):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, (3, 2, 30), 5, f_total, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_halo_window_buffer_is_not_reused_across_dtypes():
    """The window buffer is cached between steps; a float64 step after a complex128 step (bench.py times both) must get
    its own buffer instead of silently writing real rows into a complex window."""
    import torch
    from morfem_b200 import dist as mfd
    plan = mfd.build_halo_plan(0, 1, 10, [(0, 10)])
    qc = torch.randn(10, 3, dtype=torch.complex128)
    win_c = mfd.exchange_halo(qc, plan)
    qr = torch.randn(10, 3, dtype=torch.float64)
    win_r = mfd.exchange_halo(qr, plan, out=win_c)
    assert win_r.dtype == torch.float64 and torch.equal(win_r, qr)
    assert torch.equal(mfd.exchange_halo(qc, plan, out=win_c), qc)


def test_slab_halo_rows_modes():
    """H = deepest reach of any window into a neighbouring block; None (point-to-point needed) when a window reaches past a
    neighbour's slab or a rank owns fewer than H rows."""
    owned = mfd.owner_ranges(100, 4)                       # 25 rows each
    assert mfd.slab_halo_rows([(0, 25), (25, 50), (50, 75), (75, 100)], owned) == 0
    assert mfd.slab_halo_rows([(0, 30), (20, 55), (45, 80), (70, 100)], owned) == 5
    assert mfd.slab_halo_rows([(0, 60), (25, 50), (50, 75), (75, 100)], owned) is None      # rank 0 needs rows of rank 2: 35 > 25 owned
    owned = mfd.owner_ranges(100, 2)
    # rank 0's window reaches 30 rows into rank 1's block: one slab of 30 rows covers it
    assert mfd.slab_halo_rows([(0, 80), (40, 100)], owned) == 30
    # a window that skips over a whole neighbouring block needs rows from the middle of a far block
    owned = mfd.owner_ranges(90, 3)
    assert mfd.slab_halo_rows([(0, 70), (30, 60), (60, 90)], owned) is None
