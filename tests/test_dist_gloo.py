"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: agent sharding keeps batch indexing, the flat-buffer
all-reduce of per-shard gradients equals the full-batch gradient, masked means combine exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from ananke_abm_b200 import dist as abd
        from oracle import models_oracle as mo
        from oracle import torchdiffeq_oracle as tdq

        torch.set_num_threads(1)
        torch.manual_seed(0)
        B, T = 11, 4
        model = mo.OracleModeSep(8)
        g = torch.Generator().manual_seed(1)
        home = torch.randint(0, 8, (B,), generator=g)
        work = torch.randint(0, 8, (B,), generator=g)
        traits = torch.rand(B, 2, generator=g)
        mask = torch.rand(B, T, generator=g) > 0.4
        t = torch.linspace(0, 2.0, T)

        def loss_terms(h, w, tr, m):
            y0 = model.initial_state(h, w, tr)
            yp = tdq.odeint(model.rhs, y0, t, method="rk4").permute(1, 0, 2)
            per = (yp[:, :, :64] ** 2).sum(-1)
            return per[m].sum(), m.sum()

        # full batch, single process semantics: mean over masked elements
        s, c = loss_terms(home, work, traits, mask)
        (s / c).backward()
        full = abd.flatten_grads(model.parameters()).clone()
        full_loss = (s / c).detach()
        model.zero_grad()

        lo, hi = abd.shard_bounds(B, rank, world)
        sh = abd.shard_agents([home, work, traits, mask], rank, world)
        assert sh[0].shape[0] == hi - lo and torch.equal(sh[0], home[lo:hi])
        s, c = loss_terms(*sh)
        c_glob = c.clone().float()
        dist.all_reduce(c_glob)
        (s / c_glob).backward()                       # pre-scaled by the GLOBAL mask count
        flat = abd.allreduce_gradients(model.parameters())
        loss = abd.global_mean(s.detach(), c)
        ok_g = torch.allclose(flat, full, rtol=1e-5, atol=1e-7)
        ok_l = torch.allclose(loss, full_loss, rtol=1e-6)
        ok_p = all(torch.allclose(p.grad.reshape(-1), full[o:o + p.numel()], rtol=1e-5, atol=1e-7)
                   for p, o in zip(model.parameters(), _offsets(model)))
        q.put((rank, bool(ok_g), bool(ok_l), bool(ok_p), (lo, hi)))
    finally:
        dist.destroy_process_group()


def _offsets(model):
    off = 0
    for p in model.parameters():
        yield off
        off += p.numel()


def test_shard_bounds_cover_batch_without_overlap():
    from ananke_abm_b200.dist import shard_bounds
    for B in (1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gradient_allreduce_equals_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, ok_g, ok_l, ok_p, span in res:
        assert ok_g and ok_l and ok_p, (rank, ok_g, ok_l, ok_p)
    assert sorted(r[4] for r in res) == [(0, 6), (6, 11)]


def _worker_adjoint(rank, world, port, q):
    """configs[4] sharding: every rank runs the tensor-core continuous adjoint (adjoint_tc.py host algebra on the float64 stand-in
    engine of tests/test_adjoint_tc_host.py) on its block of agents with step_size; the all-reduced parameter adjoints and the
    concatenated dL/dy0 must equal the single-process solve over the whole batch -- no collective inside the solve (fixed grid)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        import numpy as np
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        sys.path.insert(0, os.path.join(root, "tests"))
        from ananke_abm_b200 import adjoint_tc, dist as abd
        import test_adjoint_tc_host as th

        torch.set_num_threads(1)
        B = 9
        oracle, y0 = th._setup(B)
        func = oracle.odefunc
        t = [0.0, 1.0]
        step = 0.25

        def solve(y0_part, n_total):
            eng = th.FakeEngine(func, 64, 32)
            rows = adjoint_tc.rk4_forward_rows(eng, y0_part, t, step, lay=th.FakeLayout, np_dtype=np.float64)
            grad_rows = 2.0 * rows / (rows.shape[0] * n_total * rows.shape[2])      # d/dy of mean(y^2) over the WHOLE batch
            return adjoint_tc.rk4_continuous_adjoint(eng, t, rows, grad_rows, step, lay=th.FakeLayout, np_dtype=np.float64, fused=True)

        gy_full, gw_full = solve(y0, B)
        lo, hi = abd.shard_bounds(B, rank, world)
        gy, gw = solve(y0[lo:hi].contiguous(), B)
        dist.all_reduce(gw)                                      # the ONE collective of a training step (dist.allreduce_gradients)
        ok_w = torch.allclose(gw, gw_full, rtol=1e-9, atol=1e-14)
        ok_y = torch.allclose(gy, gy_full[lo:hi], rtol=1e-9, atol=1e-14)
        q.put((rank, bool(ok_w), bool(ok_y), float(gw_full.abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_two_rank_sharded_continuous_adjoint_equals_the_full_batch_solve():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_adjoint, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=200) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, ok_w, ok_y, scale in res:
        assert ok_w and ok_y and scale > 0, (rank, ok_w, ok_y, scale)
