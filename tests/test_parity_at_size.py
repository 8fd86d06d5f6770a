"""-m gpu: CUDA path vs the CPU oracle AT BASELINE.json's sizes (configs[2] at 128^3 and 256^3, configs[4] for three
time steps of the moving sphere at 128^3, configs[1] P2 on 1024^2 triangles), and the solve-and-compare test the
north star names ("solution L2 error matching").

The oracle is serial C++; at 256^3 one process would need ~50 s for the 100.7 M cells, so the mesh is cut into
overlapping z-slabs (2 extra cell layers on each side: a matrix row of plane z depends on the cell layers z-2..z+1
through the ghost-penalty facets) and every slab runs the oracle pipeline of demo_poisson.py:156-201 in its own
forked process on the GLOBAL vertex coordinates and level-set values (copied back from the device, so both sides see
bit-identical inputs).  A slab reports, in global numbering, the rows / cells / facets of the planes and layers it
owns.  Bars (north star): lists and CSR pattern bit-exact, ||A - A_ref||_F / ||A_ref||_F < 1e-11, same for b,
quadrature weight sums to 1e-12 relative.
"""
import math
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_G = {}  # inherited by the forked oracle workers (no pickling of the global arrays)


def _slab_worker(args):
    lo, hi, order, gamma, gamma_g, f_value, g_value = args
    from cutfemx_b200 import mesh as M
    from oracle import pipeline

    shape, p0, p1, x, phi = _G["shape"], _G["p0"], _G["p1"], _G["x"], _G["phi"]
    tdim = len(shape)
    n_ax = shape[-1]
    L0, L1 = max(0, lo - 2), min(n_ax, hi + 2)
    vpp = int(np.prod([s + 1 for s in shape[:-1]]))
    cpl = int(np.prod(shape[:-1])) * (6 if tdim == 3 else 2)
    nslot = 12 if tdim == 3 else 3
    lshape = list(shape[:-1]) + [L1 - L0]
    h = (p1[-1] - p0[-1]) / n_ax
    q0, q1 = list(p0), list(p1)
    q0[-1], q1[-1] = p0[-1] + L0 * h, p0[-1] + L1 * h
    mesh = M.create_box(*lshape, q0, q1) if tdim == 3 else M.create_rectangle(*lshape, q0, q1)
    v0, v1 = L0 * vpp, (L1 + 1) * vpp
    mesh.x = np.ascontiguousarray(x[v0:v1])  # the global coordinates, bit for bit
    V = M.FunctionSpace(mesh, 1, mesh.x_dofmap, mesh.num_nodes, mesh.num_nodes, 1, mesh.x)
    out = pipeline.run_pipeline(mesh, V.dofmap, np.ascontiguousarray(phi[v0:v1]), V, order=order, gamma=gamma,
                                gamma_g=gamma_g, f_value=f_value, g_value=g_value)
    # owned planes [zlo, zhi) (the last slab also owns the top plane), owned cell layers [lo, hi)
    zlo, zhi = lo, (hi + 1 if hi == n_ax else hi)
    r0, r1 = (zlo - L0) * vpp, (zhi - L0) * vpp
    rp = out["row_ptr"]
    e0, e1 = int(rp[r0]), int(rp[r1])
    c0, c1 = (lo - L0) * cpl, (hi - L0) * cpl

    def cells_owned(a):
        a = np.asarray(a)
        return (a[(a >= c0) & (a < c1)] + L0 * cpl).astype(np.int64)

    gh = np.asarray(out["ghost"]).astype(np.int64)
    base = gh // nslot
    gh = gh[(base >= r0) & (base < r1)] + nslot * v0
    rv, ri = out["rv"], out["ri"]

    def wsum(r):
        pm = np.asarray(r.parent_map)
        sel = np.repeat((pm >= c0) & (pm < c1), np.diff(r.offsets))
        return float(np.sum(r.weights[sel])), int(sel.sum())

    return dict(row0=zlo * vpp, row_ptr=(rp[r0:r1 + 1] - e0).astype(np.int64),
                cols=(out["cols"][e0:e1].astype(np.int64) + v0), vals=out["vals"][e0:e1], b=out["b"][r0:r1],
                inside=cells_owned(out["inside"]), cut=cells_owned(out["cut"]), ghost=gh, wv=wsum(rv), wi=wsum(ri),
                seconds=out["total_s"])


def oracle_slabs(shape, p0, p1, x, phi, order, gamma=40.0, gamma_g=0.1, f_value=1.0, g_value=2.5, nparts=None):
    """Run the oracle pipeline on overlapping slabs; returns the per-slab dicts in ascending order."""
    import oracle

    oracle.build()
    n_ax = shape[-1]
    nparts = nparts or max(1, min(os.cpu_count() or 1, 16, n_ax // 8))
    _G.update(shape=tuple(shape), p0=tuple(p0), p1=tuple(p1), x=x, phi=phi)
    jobs = [((n_ax * p) // nparts, (n_ax * (p + 1)) // nparts, order, gamma, gamma_g, f_value, g_value)
            for p in range(nparts)]
    with mp.get_context("fork").Pool(nparts) as pool:
        return pool.map(_slab_worker, jobs)


def compare_with_slabs(slabs, rp, cols, vals, b, inside, cut, ghost, wv, wi, npv, npi):
    """GPU arrays (host copies, global numbering) against the slab results."""
    num2 = den2 = bnum2 = bden2 = 0.0
    for s in slabs:
        r0 = s["row0"]
        nr = s["row_ptr"].size - 1
        e0, e1 = int(rp[r0]), int(rp[r0 + nr])
        assert np.array_equal(rp[r0:r0 + nr + 1] - e0, s["row_ptr"]), "CSR row pointers differ from the oracle"
        assert np.array_equal(cols[e0:e1].astype(np.int64), s["cols"]), "CSR columns differ from the oracle"
        d = vals[e0:e1] - s["vals"]
        num2 += float(d @ d)
        den2 += float(s["vals"] @ s["vals"])
        db = b[r0:r0 + nr] - s["b"]
        bnum2 += float(db @ db)
        bden2 += float(s["b"] @ s["b"])
    for name, g in (("inside", inside), ("cut", cut), ("ghost", ghost)):
        ref = np.concatenate([s[name] for s in slabs])
        assert np.array_equal(np.asarray(g).astype(np.int64), ref), f"{name} list differs from the oracle"
    ea, eb = math.sqrt(num2 / den2), math.sqrt(bnum2 / bden2)
    assert ea < 1e-11 and eb < 1e-11, (ea, eb)
    wv_ref, wi_ref = sum(s["wv"][0] for s in slabs), sum(s["wi"][0] for s in slabs)
    assert abs(wv - wv_ref) <= 1e-12 * wv_ref and abs(wi - wi_ref) <= 1e-12 * wi_ref, (wv, wv_ref, wi, wi_ref)
    assert npv == sum(s["wv"][1] for s in slabs) and npi == sum(s["wi"][1] for s in slabs)
    return ea, eb


def _gpu_problem(shape, p0, p1, kind, prm, order, g_value=2.5):
    from cutfemx_b200 import demo_poisson as dp
    from cutfemx_b200.mesh import Function, FunctionSpace

    mesh = dp.device_mesh(0, list(shape), list(p0), list(p1))
    vals = dp.device_level_set(mesh, kind, prm)
    nn = int(mesh.x.shape[0])
    V = FunctionSpace(mesh, 1, mesh.x_dofmap, nn, nn, 1, None)
    phi = Function(V, "phi", vals)
    prob = dp.CutPoisson(mesh, phi, V, order=order, g_value=g_value)
    return mesh, V, phi, prob


def _gpu_step_arrays(prob):
    prob.step(keep=True)
    t = prob.last
    A = prob.A
    import importlib

    cutm = importlib.import_module("cutfemx_b200.cut")
    out = dict(rp=A.indptr, cols=A.indices, vals=A.data, b=prob.b.cpu().numpy(), inside=t["inside"].numpy(),
               ghost=t["ghost"].numpy(), cut=cutm.locate_entities(prob.cut_data, "phi=0"),
               wv=float(np.sum(t["rv"].weights)), npv=t["rv"].total_points,
               wi=float(np.sum(t["ri"].weights)), npi=t["ri"].total_points)
    prob.release_step()
    return out


@pytest.mark.parametrize("n", [128, 256])
def test_c3_sphere_against_oracle(n, built_lib):
    """configs[2]: sphere R = 0.35 on n^3 Kuhn tetrahedra, order 4, Nitsche + ghost penalty."""
    shape, p0, p1 = (n, n, n), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    mesh, V, phi, prob = _gpu_problem(shape, p0, p1, "sphere", (0.5, 0.5, 0.5, 0.35, 0.0), 4)
    g = _gpu_step_arrays(prob)
    slabs = oracle_slabs(shape, p0, p1, mesh.x.cpu().numpy(), phi.x.array.cpu().numpy(), 4)
    ea, eb = compare_with_slabs(slabs, g["rp"], g["cols"], g["vals"], g["b"], g["inside"], g["cut"], g["ghost"],
                                g["wv"], g["wi"], g["npv"], g["npi"])
    if n == 256:
        assert g["cut"].size == 691620
    print(f"C3 n={n}: |A-A_ref|/|A_ref| = {ea:.2e}, |b-b_ref|/|b_ref| = {eb:.2e}, oracle "
          f"{max(s['seconds'] for s in slabs):.1f} s on {len(slabs)} processes")


def test_c5_moving_sphere_three_steps(built_lib):
    """configs[4]: moving sphere R = 0.25 on 128^3, order 2 -- re-cut / regenerate / reassemble, steps 0, 50, 99."""
    from cutfemx_b200 import demo_poisson as dp

    n = 128
    shape, p0, p1 = (n, n, n), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    mesh, V, phi, prob = _gpu_problem(shape, p0, p1, "sphere", (0.3, 0.5, 0.5, 0.25, 0.0), 2)
    x = mesh.x.cpu().numpy()
    for t in (0, 50, 99):
        cx = 0.3 + 0.4 * t / 99
        dp.device_level_set(mesh, "sphere", (cx, 0.5, 0.5, 0.25, 0.0), out=phi.x.array)
        g = _gpu_step_arrays(prob)
        slabs = oracle_slabs(shape, p0, p1, x, phi.x.array.cpu().numpy(), 2)
        compare_with_slabs(slabs, g["rp"], g["cols"], g["vals"], g["b"], g["inside"], g["cut"], g["ghost"], g["wv"],
                           g["wi"], g["npv"], g["npi"])


def test_c2_p2_circle_1024_against_oracle(built_lib):
    """configs[1] at 1024^2: P2 u / P1 level set on right-diagonal triangles, Nitsche + P2 ghost penalty; the whole
    matrix against one serial oracle process (2.1 M cells)."""
    import torch

    import oracle as O
    from cutfemx_b200 import mesh as M
    from cutfemx_b200 import parallel as P

    n = 1024
    pipe = P.RankPipeline([n, n], [-1.0, -1.0], [1.0, 1.0], 1, 0, 0, "sphere", (0.0, 0.0, 0.0, 0.5, 0.0), order=4,
                          degree=2, g_value=2.5)
    prob = pipe.prob
    g = _gpu_step_arrays(prob)
    hm = M.create_rectangle(n, n, (-1.0, -1.0), (1.0, 1.0))
    hm.x = pipe.mesh.x.cpu().numpy()
    phi = pipe.phi.x.array.cpu().numpy()
    V1 = M.FunctionSpace(hm, 1, hm.x_dofmap, hm.num_nodes, hm.num_nodes, 1, hm.x)
    V2 = M.FunctionSpace(hm, 2, pipe.V.dofmap.cpu().numpy(), pipe.V.num_dofs, pipe.V.num_dofs, 1, None)
    from oracle import pipeline

    ref = pipeline.run_pipeline(hm, V1.dofmap, phi, V2, order=4, g_value=2.5)
    assert np.array_equal(g["inside"], ref["inside"]) and np.array_equal(g["ghost"], ref["ghost"])
    assert np.array_equal(g["rp"], ref["row_ptr"]) and np.array_equal(g["cols"], ref["cols"])
    ea = np.linalg.norm(g["vals"] - ref["vals"]) / np.linalg.norm(ref["vals"])
    eb = np.linalg.norm(g["b"] - ref["b"]) / np.linalg.norm(ref["b"])
    assert ea < 1e-11 and eb < 1e-11, (ea, eb)
    assert abs(g["wv"] - ref["rv"].weights.sum()) <= 1e-12 * ref["rv"].weights.sum()


@pytest.mark.parametrize("kind,n,tol", [("circle", 64, 2e-3), ("sphere", 24, 2e-2)])
def test_solution_l2_error_matches_oracle(kind, n, tol, built_lib):
    """North star: "solution L2 error matching".  -Laplace(u) = f in the disc / ball of radius R, u = g on the
    interface (Nitsche), ghost penalty, inactive dofs constrained to 0 (deactivate_outside) -- configs[0] is the
    circle case.  The system assembled on the GPU and the one assembled by the oracle are solved with the same
    sparse direct solver; the solutions agree to 1e-9 relative, and the L2 errors against the exact solution
    u = g + f (R^2 - r^2) / (2 d), measured with the quadratic functional of demo_poisson.py:213 on the cut domain
    (CFX_K_SQUARE_FN on the GPU, the oracle's restatement on the CPU), agree to 1e-8 relative."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    import cutfemx_b200 as cfx
    import oracle as O
    from util import GpuRun, OracleRun, make_problem

    f_value, g_value = 1.0, 0.25
    mesh, Vphi, phi, V = make_problem(kind, n)
    d = mesh.tdim
    R = 0.5 if kind == "circle" else 0.35
    cx = (0.0, 0.0, 0.0) if kind == "circle" else (0.5, 0.5, 0.5)
    X = V.dof_coords
    r2 = sum((X[:, k] - cx[k]) ** 2 for k in range(3 if d == 3 else 2))
    u_exact = g_value + f_value * (R * R - r2) / (2.0 * d)

    ref = OracleRun(mesh, Vphi, phi, V, order=4, g_value=g_value, f_value=f_value)
    gpu = GpuRun(mesh, Vphi, phi, V, order=4, g_value=g_value, f_value=f_value)
    # deactivate_outside on both sides (fem/deactivate.h:387-418)
    _, inactive_ref = O.active_domain(V, [ref.inside, ref.rv.parent_map, ref.ri.parent_map], ref.rows4)
    vals_ref, b_ref = ref.vals.copy(), ref.b.copy()
    O.deactivate_outside(ref.row_ptr, ref.cols, vals_ref, inactive_ref, 1.0, b_ref, 0.0)
    dom = cfx.fem.active_domain(gpu.a)
    assert np.array_equal(dom.inactive_dofs, inactive_ref)
    import torch

    bg = torch.from_numpy(gpu.b.copy()).cuda()
    cfx.fem.deactivate_outside(gpu.A, bg, dom, diagonal=1.0, rhs_value=0.0)
    Ag = sp.csr_matrix((gpu.A.data, gpu.A.indices, gpu.A.indptr), shape=gpu.A.shape)
    Ar = sp.csr_matrix((vals_ref, ref.cols, ref.row_ptr), shape=gpu.A.shape)
    ug = spla.spsolve(Ag.tocsc(), bg.cpu().numpy())
    ur = spla.spsolve(Ar.tocsc(), b_ref)
    assert np.linalg.norm(ug - ur) <= 1e-9 * np.linalg.norm(ur)

    def l2_error_gpu(u):
        M0 = cfx.fem.CutForm(V, 0).add_cell_integral("square_fn", gpu.inside, gpu.rv, (1.0,)).set_coefficient(u - u_exact)
        return math.sqrt(cfx.fem.assemble_scalar(M0))

    def l2_error_ref(u):
        val = np.zeros(1)
        with O.coefficient(u - u_exact):
            O.assemble_cells(V, "square_fn", val, ref.inside, ref.rv, (1.0,))
        return math.sqrt(val[0])

    eg, er = l2_error_gpu(ug), l2_error_ref(ur)
    assert abs(eg - er) <= 1e-8 * er, (eg, er)
    assert er < tol, er  # and the discrete solution is the right one


@pytest.mark.parametrize("n", [14])
def test_configs3_elasticity_on_the_device_generated_p2_vector_space(n, built_lib):
    """BASELINE configs[3] (torus R = 0.3, r = 0.12; linear elasticity on a P2 VECTOR space, vector Nitsche terms on
    the interface, componentwise ghost penalty; demo_elasticity.py:213-238) through the path bench.py --workload C4
    times -- device-generated mesh, device-generated P2-tetrahedron dofmap (cfx_meshgen_p2_tet_dofmap), standard
    cells evaluated on the fly by the blocked row gather, cut-cell tensors by the warp-per-rule kernel -- against the
    oracle's run_elasticity_pipeline on the same dofmap.  Bars: pattern bit-exact, 3 x 3 blocks and the load vector
    to 1e-11 relative Frobenius; the matrix is symmetric and the load sums to f x wet volume per component."""
    import torch

    from cutfemx_b200 import mesh as M
    from cutfemx_b200 import parallel as P
    from oracle import pipeline

    pipe = P.RankPipeline([n] * 3, [0.0] * 3, [1.0] * 3, 1, 0, 0, "torus", (0.5, 0.5, 0.5, 0.3, 0.12, 0.0), order=4,
                          degree=2, problem="elasticity", bs=3)
    prob = pipe.prob
    pipe.xplan = None
    prob.persistent = False
    prob.step(keep=True)
    torch.cuda.synchronize()
    A, b = prob.A, prob.b.cpu().numpy()
    hm = M.create_box(n, n, n)
    assert np.array_equal(hm.x_dofmap, pipe.mesh.x_dofmap.cpu().numpy())       # same numbering as the host generator
    np.testing.assert_array_equal(hm.x, pipe.mesh.x.cpu().numpy())
    dm = pipe.V.dofmap.cpu().numpy()
    assert dm.shape == (hm.num_cells, 10) and np.array_equal(dm[:, :4], hm.x_dofmap)
    # the edge dofs are a consistent numbering: equal to the host generator's up to a renumbering of the edge dofs
    Vh = M.functionspace(hm, 2)
    _, first = np.unique(Vh.dofmap[:, 4:].ravel(), return_index=True)
    assert np.unique(dm[:, 4:].ravel()).size == first.size
    assert np.array_equal(np.unique(dm[:, 4:].ravel()[first], return_counts=True)[1], np.ones(first.size, dtype=np.int64))
    V = M.FunctionSpace(hm, 2, dm, pipe.V.num_dofs, pipe.V.num_dofs, 3, None)
    V1 = M.functionspace(hm, 1)
    phi = pipe.phi.x.array.cpu().numpy()
    ref = pipeline.run_elasticity_pipeline(hm, V1.dofmap, phi, V, order=4)
    assert np.array_equal(A.indptr, ref["row_ptr"]) and np.array_equal(A.indices, ref["cols"])
    ea = np.linalg.norm(A.data - ref["vals"]) / np.linalg.norm(ref["vals"])
    eb = np.linalg.norm(b - ref["b"]) / np.linalg.norm(ref["b"])
    assert ea < 1e-11 and eb < 1e-11, (ea, eb)
    Ms = A.to_scipy()
    assert abs(Ms - Ms.T).max() <= 1e-11 * abs(Ms).max()
    wet = ref["rv"].weights.sum() + ref["inside"].size / (6.0 * n ** 3)
    np.testing.assert_allclose(b.reshape(-1, 3).sum(axis=0), np.array([0.0, 0.0, -1.0]) * wet, rtol=0, atol=1e-12 * wet)
    prob.release_step()
