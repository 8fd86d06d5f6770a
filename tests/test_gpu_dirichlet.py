"""-m gpu: Dirichlet conditions on the assembled system against the oracle's
restatement of assemble_matrix(bcs) / set_diagonal / apply_lifting / set_bc
(assembler.h:643-683,745-787; assemble_matrix_impl.h:146-185,537-603; assemble_vector_impl.h:383-564;
flow of demo_elasticity.py:78-84).  Bars: matrix and vector to 1e-11 relative Frobenius."""
import numpy as np
import pytest

import oracle as O
from cutfemx_b200 import mesh as M
from util import make_problem

pytestmark = pytest.mark.gpu
MU, LAM, GAMMA_G = 384.6, 576.9, 0.05


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _boundary_dofs(V, axis, value, comps):
    """blocked indices bs*dof + k of the dofs on the plane x[axis] == value (locate_dofs_topological on a
    boundary facet set, demo_elasticity.py:246-259)."""
    on = np.nonzero(np.isclose(V.dof_coords[:, axis], value))[0]
    return np.sort(np.concatenate([on * V.bs + k for k in comps])).astype(np.int32)


CASES = [("circle", 12, 1, 1), ("circle", 8, 2, 1), ("sphere", 6, 1, 1), ("circle", 10, 1, 2), ("sphere", 5, 1, 3),
         ("circle", 6, 2, 2)]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"{c[0]}{c[1]}-P{c[2]}-bs{c[3]}")
def problem(request, built_lib):
    import cutfemx_b200 as cfx

    kind, n, deg, bs = request.param
    mesh, Vphi, phi, _ = make_problem(kind, n, 1)
    V = M.functionspace(mesh, deg, bs=bs, permute_seed=5)
    vals = phi.x.array
    dom = O.classify(Vphi.dofmap, vals)
    inside, cut = O.locate(dom, "phi<0"), O.locate(dom, "phi=0")
    rv = O.runtime_quadrature(mesh, Vphi.dofmap, vals, dom, "<", 4)
    ghost = O.ghost_penalty_facets(mesh, cut, inside)
    rows4 = O.facet_rows(mesh, ghost)
    rp, cols = O.sparsity(V, np.concatenate([inside, rv.parent_map]), rows4)
    cell_kernel, cc = ("laplace", (1.3,)) if bs == 1 else ("elasticity", (MU, LAM))
    gp = (0.1,) if bs == 1 else (GAMMA_G * (2 * MU + LAM),)

    def oracle_matrix(out):
        O.assemble_cells(V, cell_kernel, out, inside, rv, cc, rp, cols)
        O.assemble_interior_facets(V, "ghost_grad_jump", out, rows4, gp, rp, cols)
        return out

    # conditions: a plane through the cut domain clamps every component, a second plane one component
    x0p = float(np.median(V.dof_coords[:, 0]))
    y0p = float(np.unique(V.dof_coords[:, 1])[len(np.unique(V.dof_coords[:, 1])) // 3])
    rng = np.random.default_rng(7)
    g1 = rng.standard_normal(V.num_dofs * bs)
    bcs = [cfx.fem.dirichletbc(0.25, _boundary_dofs(V, 0, x0p, range(bs)), V),
           cfx.fem.dirichletbc(g1, _boundary_dofs(V, 1, y0p, [0]), V)]
    markers, values = cfx.fem._bc_arrays(V, bcs)
    assert markers.sum() > 0

    cd = cfx.cut(phi)
    g_inside = cfx.locate_entities(cd, "phi<0")
    g_rv = cfx.runtime_quadrature(cd, "phi<0", 4)
    a = cfx.fem.CutForm(V, 2).add_cell_integral(cell_kernel, g_inside, g_rv, cc)
    a.add_interior_facet_integral("ghost_grad_jump", facets=cfx.ghost_penalty_facets(cd, "phi<0"), constants=gp)
    return dict(cfx=cfx, V=V, bs=bs, rp=rp, cols=cols, oracle_matrix=oracle_matrix, bcs=bcs, markers=markers,
                values=values, a=a)


def test_matrix_with_bcs(problem):
    p = problem
    cfx, bs = p["cfx"], p["bs"]
    ref = np.zeros(p["cols"].size * bs * bs)
    with O.dirichlet("matrix", p["markers"], p["markers"]):
        p["oracle_matrix"](ref)
    for bc in p["bcs"]:
        O.set_diagonal(p["rp"], p["cols"], ref, bc.owned_dofs(), 1.0, bs)
    A = cfx.fem.assemble_matrix(p["a"], bcs=p["bcs"])
    assert np.array_equal(A.indptr, p["rp"]) and np.array_equal(A.indices, p["cols"])
    assert rel(A.data, ref) < 1e-11
    # Dirichlet rows and columns are exactly the identity pattern
    Ms = A.to_scipy()
    d = np.nonzero(p["markers"])[0]
    sub = Ms[d]
    assert np.all(sub.data[sub.indices != np.repeat(d, np.diff(sub.indptr))] == 0.0)
    assert np.all(Ms.diagonal()[d] == 1.0)
    assert abs(Ms - Ms.T).max() <= 1e-12 * abs(Ms).max()


def test_accumulating_assembly_keeps_previous_values(problem):
    """assemble_matrix ADDS (assembler.h:596-597): with bcs the masked contributions add nothing, so a second
    assembly into the same matrix doubles every entry except the diagonal that insert_diagonal SETS."""
    p = problem
    cfx, bs = p["cfx"], p["bs"]
    ref = np.zeros(p["cols"].size * bs * bs)
    with O.dirichlet("matrix", p["markers"], p["markers"]):
        p["oracle_matrix"](ref)
        for bc in p["bcs"]:
            O.set_diagonal(p["rp"], p["cols"], ref, bc.owned_dofs(), 1.0, bs)
        p["oracle_matrix"](ref)
    for bc in p["bcs"]:
        O.set_diagonal(p["rp"], p["cols"], ref, bc.owned_dofs(), 1.0, bs)
    A = cfx.fem.assemble_matrix(p["a"], bcs=p["bcs"])
    cfx.fem.assemble_matrix(p["a"], A, bcs=p["bcs"])
    assert rel(A.data, ref) < 1e-11


def test_lifting_and_set_bc(problem):
    p = problem
    cfx, V, bs = p["cfx"], p["V"], p["bs"]
    n = V.num_dofs * bs
    rng = np.random.default_rng(3)
    for x0, alpha in ((None, 1.0), (rng.standard_normal(n), -0.75)):
        b0 = rng.standard_normal(n)
        bref = b0.copy()
        with O.dirichlet("lifting", None, p["markers"], p["values"], x0, alpha, bref):
            p["oracle_matrix"](np.zeros(p["cols"].size * bs * bs))
        b = b0.copy()
        A = cfx.fem.create_matrix(p["a"])
        before = A.data.copy()
        cfx.fem.apply_lifting(b, [p["a"]], [p["bcs"]], None if x0 is None else [x0], alpha, A=[A])
        assert rel(b - b0, bref - b0) < 1e-11 and np.linalg.norm(bref - b0) > 0
        A._cache.clear()
        assert np.array_equal(A.data, before)  # the matrix values are left as they were
        for bc in p["bcs"]:
            O.set_bc(bref, bc.dofs, bc.g, x0, alpha)
        cfx.fem.set_bc(b, p["bcs"], x0, alpha)
        assert rel(b, bref) < 1e-11
        d = np.nonzero(p["markers"])[0]
        assert np.array_equal(b[d], bref[d])


def test_constrained_solve_reproduces_a_linear_field(built_lib):
    """End to end, the flow of demo_elasticity.py:78-95 on a Poisson problem: every active dof except an interior
    patch carries the values of a linear function; Laplace and the gradient-jump penalty both vanish on linear
    fields against interior test functions, so the constrained solve must return the linear field on the patch."""
    import scipy.sparse.linalg as spla

    import cutfemx_b200 as cfx

    mesh, Vphi, phi, _ = make_problem("circle", 10, 1)
    V = M.functionspace(mesh, 1)
    cd = cfx.cut(phi)
    inside = cfx.locate_entities(cd, "phi<0")
    rv = cfx.runtime_quadrature(cd, "phi<0", 2)
    a = cfx.fem.CutForm(V, 2).add_cell_integral("laplace", inside, rv, (1.0,))
    a.add_interior_facet_integral("ghost_grad_jump", facets=cfx.ghost_penalty_facets(cd, "phi<0"), constants=(0.1,))
    L = cfx.fem.CutForm(V, 1).add_cell_integral("source", inside, rv, (0.0,))
    ad = cfx.fem.active_domain(a)
    u_lin = 0.3 + 1.7 * V.dof_coords[:, 0] - 0.9 * V.dof_coords[:, 1]
    active = np.setdiff1d(np.arange(V.num_dofs), ad.inactive_dofs)
    r = np.linalg.norm(V.dof_coords[:, :2], axis=1)
    free = active[r[active] < 0.25]
    assert free.size > 3
    bc = cfx.fem.dirichletbc(u_lin, np.setdiff1d(active, free), V)
    A = cfx.fem.assemble_matrix(a, bcs=[bc])
    b = cfx.fem.assemble_vector(L)
    cfx.fem.apply_lifting(b, [a], [[bc]], A=[A])
    cfx.fem.set_bc(b, [bc])
    cfx.fem.deactivate_outside(A, b, ad)
    u = spla.spsolve(A.to_scipy().tocsc(), b)
    np.testing.assert_allclose(u[active], u_lin[active], rtol=0, atol=1e-11)


def test_system_with_bcs_in_one_assembly(problem):
    """cfx_assemble_system_bc == assemble_matrix(bcs) + assemble_vector + apply_lifting + set_bc, the sequence of
    demo_elasticity.py:78-84, with one assembly of the unconstrained system."""
    import torch

    p = problem
    cfx, V, bs = p["cfx"], p["V"], p["bs"]
    n = V.num_dofs * bs
    rng = np.random.default_rng(11)
    x0, alpha = rng.standard_normal(n), 0.8
    L = cfx.fem.CutForm(V, 1)  # zero load, as in demo_elasticity.py:238
    # the reference sequence, call by call
    A1 = cfx.fem.assemble_matrix(p["a"], bcs=p["bcs"])
    b1 = cfx.fem.assemble_vector(L)
    cfx.fem.apply_lifting(b1, [p["a"]], [p["bcs"]], [x0], alpha, A=[A1])
    cfx.fem.set_bc(b1, p["bcs"], x0, alpha)
    # one call
    A2 = cfx.fem.create_matrix(p["a"])
    b2 = torch.zeros(n, dtype=torch.float64, device="cuda:0")
    cfx.fem.assemble_system_bc(p["a"], A2, L, b2, p["bcs"], x0, alpha)
    assert np.array_equal(A2.indptr, A1.indptr) and np.array_equal(A2.indices, A1.indices)
    assert rel(A2.data, A1.data) < 1e-13
    assert rel(b2.cpu().numpy(), b1) < 1e-11
