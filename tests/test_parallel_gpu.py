"""-m gpu: the N>1 device path -- partition on the device, ghost-row pattern (cfx_create_sparsity_rows),
inserted entries (cfx_form_insert_pattern_entries), positions, pack / unpack-add kernels -- against the
SERIAL oracle assembly of the same problem.

* N ranks emulated on one GPU (one context per rank, mailbox transport): always runs;
* two processes over NCCL: runs when the box has >= 2 GPUs (gpurun --gpus 2).
"""
import os
import sys

import numpy as np
import pytest

from test_parallel_gloo import box, compare_with_serial, level_set

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(order=4, gamma=40.0, gamma_g=0.1, f_value=1.0, g_value=2.5)


def make_pipes(shape, world, devices):
    from cutfemx_b200 import parallel as P

    p0, p1 = box(len(shape))
    return [P.RankPipeline(shape, p0, p1, world, r, devices[r], level_set(len(shape)), None, **KW) for r in range(world)]


@pytest.mark.parametrize("shape,world", [((8, 8, 8), 2), ((6, 6, 9), 3), ((16, 16), 2), ((20, 20), 2), ((6, 6, 6), 1),
                                         ((12, 10, 16), 4), ((6, 6, 16), 8)])
def test_emulated_ranks_match_serial(shape, world, built_lib):
    from cutfemx_b200 import parallel as P

    pipes = make_pipes(shape, world, [0] * world)
    tr = P.LocalTransport(world)
    P.plan(pipes, tr)
    P.run_step(pipes, tr)
    parts = [p.owned_matrix_global() for p in pipes]
    compare_with_serial(parts, shape)
    # a second step (moving-domain loop: everything is rebuilt) gives bit-identical results
    vals1 = [p.prob.A.data.copy() for p in pipes]
    for p in pipes:
        p.finish_step()
    P.run_step(pipes, tr)
    for p, v in zip(pipes, vals1):
        assert np.array_equal(p.prob.A.data, v)


def _nccl_worker(rank, world, shape, port, outdir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from cutfemx_b200 import parallel as P

        p0, p1 = box(len(shape))
        pipe = P.RankPipeline(shape, p0, p1, world, rank, rank, level_set(len(shape)), None, **KW)
        tr = P.TorchDistTransport()
        P.plan([pipe], tr)
        P.run_step([pipe], tr)
        rows, cols, vals, b, off = pipe.owned_matrix_global()
        np.savez(os.path.join(outdir, f"part{rank}.npz"), rows=rows, cols=cols, vals=vals, b=b, off=off)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpus_nccl_match_serial(tmp_path, built_lib):
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    shape = (12, 12, 12)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(2, shape, port, str(tmp_path)), nprocs=2, join=True)
    parts = []
    for rank in range(2):
        d = np.load(tmp_path / f"part{rank}.npz")
        parts.append((d["rows"], d["cols"], d["vals"], d["b"], int(d["off"])))
    compare_with_serial(parts, shape)


# ----------------------------------------------------------------------------- static exchange plan (cfx_xplan_*)
@pytest.mark.parametrize("shape,world", [((8, 8, 8), 2), ((6, 6, 9), 3), ((16, 16), 2), ((12, 10, 16), 4), ((6, 6, 16), 8)])
def test_static_plan_emulated_ranks_match_serial(shape, world, built_lib):
    """Fixed-size bit / value messages, inserted entries expanded on the device, pack / unpack-add kernels; the
    transport between the ranks emulated on one GPU is a device copy of the message buffers.  Then the same step with
    deferred sizes: bit-identical."""
    from cutfemx_b200 import parallel as P

    pipes = make_pipes(shape, world, [0] * world)
    P.plan(pipes, P.LocalTransport(world), static=True)
    for p in pipes:
        p.prob.persistent = True
        p.ctx.set_deferred(False, 0.25)
    P.run_step_static(pipes)

    def parts():
        for p in pipes:
            p.prob.A._cache.clear()
        return [p.owned_matrix_global() for p in pipes]

    ref = parts()
    compare_with_serial(ref, shape)
    for p in pipes:
        p.finish_step()
    P.run_step_static(pipes)
    for p in pipes:
        p.finish_step()
        p.ctx.set_deferred(True)
    P.run_step_static(pipes)            # nothing reaches the host in this step
    for p in pipes:
        p.finish_step()
        p.ctx.check()
    for a, b in zip(parts(), ref):
        for x, y in zip(a[:4], b[:4]):
            assert np.array_equal(x, y)


def _nccl_static_worker(rank, world, shape, port, outdir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from cutfemx_b200 import parallel as P

        p0, p1 = box(len(shape))
        pipe = P.RankPipeline(shape, p0, p1, world, rank, rank, level_set(len(shape)), None, **KW)
        P.plan([pipe], P.TorchDistTransport(), static=True)
        P.init_nccl(pipe.ctx, rank, world)
        g = pipe.capture_static()       # eager, deferred and captured steps, NCCL inside the graph
        for _ in range(2):
            pipe.prob.replay()
        pipe.ctx.check()
        pipe.prob.A._cache.clear()
        rows, cols, vals, b, off = pipe.owned_matrix_global()
        np.savez(os.path.join(outdir, f"part{rank}.npz"), rows=rows, cols=cols, vals=vals, b=b, off=off,
                 nodes=g.kernel_nodes)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(8, 8, 8), (24, 24)])
def test_two_gpus_static_plan_graph_match_serial(shape, tmp_path, built_lib):
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_static_worker, args=(2, shape, port, str(tmp_path)), nprocs=2, join=True)
    parts = []
    for rank in range(2):
        d = np.load(tmp_path / f"part{rank}.npz")
        parts.append((d["rows"], d["cols"], d["vals"], d["b"], int(d["off"])))
    compare_with_serial(parts, shape)


# ----------------------------------------------------------------------------- blocked / P2 spaces on several ranks
def _scalar_coo(rows, cols, vals, bs, n):
    import scipy.sparse as sp

    i = np.arange(bs)
    r = (rows[:, None, None] * bs + i[None, :, None]) + 0 * i[None, None, :]
    c = (cols[:, None, None] * bs + i[None, None, :]) + 0 * i[None, :, None]
    A = sp.csr_matrix((vals.reshape(-1), (r.reshape(-1), c.reshape(-1))), shape=(n, n))
    assert A.nnz == r.size  # no duplicate (row, col)
    A.sort_indices()
    return A


@pytest.mark.parametrize("shape,world,degree", [((6, 6, 8), 2, 2), ((4, 4, 9), 3, 2), ((6, 6, 8), 2, 1), ((4, 4, 5), 1, 2)])
def test_static_plan_blocked_spaces_match_serial(shape, world, degree, built_lib):
    """BASELINE configs[3] on N ranks: linear elasticity on a P2 (and P1) VECTOR space, slab partition with the P2
    dofs numbered by parallel.p2_tet_slab_space, static exchange plan with COMPACT value messages (bs x bs blocks of
    the entries present this step only).  The union of the ranks' owned rows equals the one-rank device assembly
    (itself compared with the oracle in test_gpu_elasticity): pattern bit-exact, blocks and right-hand side 1e-11."""
    from cutfemx_b200 import parallel as P

    p0, p1 = box(3)
    kind, prm = "sphere", (0.5, 0.5, 0.5, 0.35, 0.0)
    kw = dict(order=4, degree=degree, problem="elasticity", bs=3)
    one = P.RankPipeline(shape, p0, p1, 1, 0, 0, kind, prm, **kw)
    one.step_static()
    A1 = one.prob.A
    n1 = one.V.num_dofs
    rows1 = np.repeat(np.arange(n1), np.diff(A1.indptr))
    cols1 = A1.indices.astype(np.int64)
    vals1 = A1.data.reshape(-1, 3, 3)
    b1 = one.prob.b.cpu().numpy().reshape(n1, 3)
    used = np.zeros(n1, dtype=bool)
    used[np.unique(one.V.dofmap.cpu().numpy())] = True
    keep = used[rows1]
    ref = _scalar_coo(rows1[keep], cols1[keep], vals1[keep], 3, 3 * n1)

    pipes = [P.RankPipeline(shape, p0, p1, world, r, 0, kind, prm, p2_numbering="blocks", **kw) for r in range(world)]
    P.plan(pipes, P.LocalTransport(world), static=True)
    for p in pipes:
        p.prob.persistent = True
        p.ctx.set_deferred(False, 0.25)
    P.run_step_static(pipes)

    def parts():
        for p in pipes:
            p.prob.A._cache.clear()
        return [p.owned_matrix_global() for p in pipes]

    def to_serial(g):
        return P.p2_tet_global_to_serial(g, shape) if degree == 2 else g

    got = parts()
    rows = np.concatenate([to_serial(q[0]) for q in got])
    cols = np.concatenate([to_serial(q[1]) for q in got])
    vals = np.concatenate([q[2] for q in got])
    keep = used[rows]
    A = _scalar_coo(rows[keep], cols[keep], vals[keep], 3, 3 * n1)
    assert np.array_equal(A.indptr, ref.indptr) and np.array_equal(A.indices, ref.indices)
    assert np.linalg.norm(A.data - ref.data) <= 1e-11 * np.linalg.norm(ref.data)
    b = np.zeros((n1, 3))
    for q in got:
        ids = to_serial(np.arange(q[4], q[4] + q[3].shape[0]))
        b[ids] = q[3]
    assert np.linalg.norm(b - b1) <= 1e-11 * np.linalg.norm(b1)
    assert np.linalg.norm(b1) > 0 and ref.nnz > 0
    # a deferred-size step (message capacities instead of exact sizes, nothing reaches the host): bit-identical
    for p in pipes:
        p.finish_step()
    P.run_step_static(pipes)
    for p in pipes:
        p.finish_step()
        p.ctx.set_deferred(True)
    P.run_step_static(pipes)
    for p in pipes:
        p.finish_step()
        p.ctx.check()
    for x, y in zip(parts(), got):
        for u, v in zip(x[:4], y[:4]):
            assert np.array_equal(u, v)


def _nccl_blocked_worker(rank, world, shape, port, outdir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from cutfemx_b200 import parallel as P

        p0, p1 = box(3)
        pipe = P.RankPipeline(shape, p0, p1, world, rank, rank, "sphere", (0.5, 0.5, 0.5, 0.35, 0.0), order=4,
                              degree=2, problem="elasticity", bs=3)
        P.plan([pipe], P.TorchDistTransport(), static=True)
        P.init_nccl(pipe.ctx, rank, world)
        g = pipe.capture_static()       # eager (exact message sizes), deferred and captured steps (capacities)
        for _ in range(2):
            pipe.prob.replay()
        pipe.ctx.check()
        pipe.prob.A._cache.clear()
        rows, cols, vals, b, off = pipe.owned_matrix_global()
        np.savez(os.path.join(outdir, f"part{rank}.npz"), rows=rows, cols=cols, vals=vals, b=b, off=off,
                 nodes=g.kernel_nodes)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpus_blocked_p2_graph_match_serial(tmp_path, built_lib):
    """configs[3] on two GPUs: the whole rank step, NCCL exchange of compact block messages included, as one CUDA
    graph; union of the owned rows == the one-rank device assembly."""
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from cutfemx_b200 import parallel as P

    shape = (6, 6, 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_blocked_worker, args=(2, shape, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = box(3)
    one = P.RankPipeline(shape, p0, p1, 1, 0, 0, "sphere", (0.5, 0.5, 0.5, 0.35, 0.0), order=4, degree=2,
                         problem="elasticity", bs=3)
    one.step_static()
    A1, n1 = one.prob.A, one.V.num_dofs
    rows1 = np.repeat(np.arange(n1), np.diff(A1.indptr))
    used = np.zeros(n1, dtype=bool)
    used[np.unique(one.V.dofmap.cpu().numpy())] = True
    keep = used[rows1]
    ref = _scalar_coo(rows1[keep], A1.indices.astype(np.int64)[keep], A1.data.reshape(-1, 3, 3)[keep], 3, 3 * n1)
    b1 = one.prob.b.cpu().numpy().reshape(n1, 3)
    rows, cols, vals, b = [], [], [], np.zeros((n1, 3))
    for rank in range(2):
        d = np.load(tmp_path / f"part{rank}.npz")
        rows.append(P.p2_tet_global_to_serial(d["rows"], shape))
        cols.append(P.p2_tet_global_to_serial(d["cols"], shape))
        vals.append(d["vals"])
        off = int(d["off"])
        b[P.p2_tet_global_to_serial(np.arange(off, off + d["b"].shape[0]), shape)] = d["b"]
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    keep = used[rows]
    A = _scalar_coo(rows[keep], cols[keep], vals[keep], 3, 3 * n1)
    assert np.array_equal(A.indptr, ref.indptr) and np.array_equal(A.indices, ref.indices)
    assert np.linalg.norm(A.data - ref.data) <= 1e-11 * np.linalg.norm(ref.data)
    assert np.linalg.norm(b - b1) <= 1e-11 * np.linalg.norm(b1)
