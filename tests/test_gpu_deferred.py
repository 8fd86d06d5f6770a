"""-m gpu: deferred-size mode and CUDA-graph replay of the whole step (include/cutfemx_b200.h "deferred sizes and CUDA
graphs"; the time loop of python/demo/demo_moving_poisson.py:69-107).

The bar is bit-exactness: a step whose sizes never reach the host, and a step replayed from a captured graph after
the level set moved, produce exactly the lists, rules, CSR pattern, matrix values and right-hand side of a fresh
eager step on the same level set -- same kernels, same order, same summation order.  A result that outgrows the
capacity of the objects being reused must surface as an error, never as a silently truncated result.
"""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cutm = importlib.import_module("cutfemx_b200.cut")


def make(shape, p0, p1, prm, order=2, degree=1):
    """Device-resident mesh, level set and problem (sphere / circle level set with parameters `prm`)."""
    import torch

    from cutfemx_b200 import demo_poisson as dp
    from cutfemx_b200.mesh import Function, FunctionSpace

    mesh = dp.device_mesh(0, list(shape), list(p0), list(p1))
    vals = dp.device_level_set(mesh, "sphere", prm)
    nn = int(mesh.x.shape[0])
    Vphi = FunctionSpace(mesh, 1, mesh.x_dofmap, nn, nn, 1, None)
    phi = Function(Vphi, "phi", vals)
    V = Vphi
    if degree == 2:  # P2 on triangles: vertex dofs + one dof per edge (= facet)
        dm = torch.cat([mesh.x_dofmap, mesh.c2f + nn], dim=1).contiguous()
        nd = nn + int(mesh.num_facets)
        V = FunctionSpace(mesh, 2, dm, nd, nd, 1, None)
    prob = dp.CutPoisson(mesh, phi, V, order=order, g_value=2.5)
    return mesh, phi, prob


def move(mesh, phi, prm):
    from cutfemx_b200 import demo_poisson as dp

    dp.device_level_set(mesh, "sphere", prm, out=phi.x.array)


def results(prob):
    """Everything a step produced, on the host (resolves deferred sizes)."""
    k = prob.last if prob.last else prob.keep
    A = prob.A
    A._cache.clear()
    for r in (k["rv"], k["ri"]):
        r._cache.clear()
    return dict(rp=A.indptr.copy(), cols=A.indices.copy(), vals=A.data.copy(), b=prob.b.cpu().numpy().copy(),
                inside=k["inside"].numpy(), ghost=k["ghost"].numpy(), rows=k["rows"].numpy(),
                wv=k["rv"].weights.copy(), pv=k["rv"].points.copy(), ov=k["rv"].offsets.copy(),
                mv=k["rv"].parent_map.copy(), wi=k["ri"].weights.copy(), pi=k["ri"].points.copy())


def assert_same(a, b):
    for key in a:
        assert a[key].shape == b[key].shape, (key, a[key].shape, b[key].shape)
        assert np.array_equal(a[key], b[key]), f"{key} differs"


CASES = [
    ("sphere16", (16, 16, 16), (0.0,) * 3, (1.0,) * 3, (0.45, 0.5, 0.5, 0.3, 0.0), (0.55, 0.48, 0.5, 0.3, 0.0), 2, 1),
    ("circle48", (48, 48), (-1.0, -1.0), (1.0, 1.0), (0.0, 0.0, 0.0, 0.5, 0.0), (0.07, -0.03, 0.0, 0.5, 0.0), 4, 1),
    ("circle32_p2", (32, 32), (-1.0, -1.0), (1.0, 1.0), (0.0, 0.0, 0.0, 0.5, 0.0), (0.05, 0.02, 0.0, 0.5, 0.0), 4, 2),
]


@pytest.mark.parametrize("name,shape,p0,p1,prm0,prm1,order,degree", CASES, ids=[c[0] for c in CASES])
def test_deferred_and_graph_steps_are_bit_identical_to_eager(name, shape, p0, p1, prm0, prm1, order, degree, built_lib):
    # eager references on their own meshes (one context per mesh)
    _, _, ref0 = make(shape, p0, p1, prm0, order, degree)
    ref0.step(keep=True)
    r0 = results(ref0)
    _, _, ref1 = make(shape, p0, p1, prm1, order, degree)
    ref1.step(keep=True)
    r1 = results(ref1)

    mesh, phi, prob = make(shape, p0, p1, prm0, order, degree)
    ctx = prob.ctx
    prob.persistent = True
    ctx.set_deferred(False, 0.25)
    prob.step()
    prob.step()                       # objects refilled in place, still eager
    assert_same(results(prob), r0)
    ctx.set_deferred(True)
    prob.step()                       # no size reaches the host in this step
    ctx.check()
    assert_same(results(prob), r0)
    move(mesh, phi, prm1)
    prob.step()                       # deferred step on a different level set: the counts change, capacities hold
    ctx.check()
    assert_same(results(prob), r1)

    # graph: capture one step, replay it for both level sets
    n0 = ctx.launch_count
    ctx.graph_begin()
    prob.step()
    g = ctx.graph_end()
    assert g.kernel_nodes > 10
    assert ctx.launch_count - n0 >= g.kernel_nodes  # captured launches are counted once at capture ...
    move(mesh, phi, prm0)
    n1 = ctx.launch_count
    g.launch()
    assert ctx.launch_count - n1 == g.kernel_nodes  # ... and once per replay
    ctx.check()
    assert_same(results(prob), r0)
    move(mesh, phi, prm1)
    g.launch()
    g.launch()                        # replaying twice changes nothing
    ctx.check()
    assert_same(results(prob), r1)
    g.free()
    # and the context still works eagerly afterwards
    ctx.set_deferred(False)
    move(mesh, phi, prm0)
    prob.step()
    assert_same(results(prob), r0)


def test_capture_helper_and_moving_loop(built_lib):
    """CutPoisson.capture() / replay(): the moving sphere of configs[4] for a few time steps, each replay compared
    with a fresh eager step."""
    shape, p0, p1 = (20, 20, 20), (0.0,) * 3, (1.0,) * 3
    mesh, phi, prob = make(shape, p0, p1, (0.3, 0.5, 0.5, 0.25, 0.0), 2)
    g = prob.capture(margin=0.3)
    assert g.kernel_nodes > 10
    for t in (0, 33, 66, 99):
        prm = (0.3 + 0.4 * t / 99, 0.5, 0.5, 0.25, 0.0)
        move(mesh, phi, prm)
        prob.replay()
        prob.ctx.check()
        _, _, ref = make(shape, p0, p1, prm, 2)
        ref.step(keep=True)
        assert_same(results(prob), results(ref))


def test_capacity_overflow_is_an_error_not_a_truncation(built_lib):
    from cutfemx_b200._lib import CfxError

    shape, p0, p1 = (24, 24, 24), (0.0,) * 3, (1.0,) * 3
    small, large = (0.5, 0.5, 0.5, 0.12, 0.0), (0.5, 0.5, 0.5, 0.45, 0.0)
    mesh, phi, prob = make(shape, p0, p1, small, 2)
    ctx = prob.ctx
    prob.persistent = True
    ctx.set_deferred(False, 0.0)
    prob.step()
    prob.step()
    ctx.set_deferred(True)
    prob.step()
    ctx.check()
    move(mesh, phi, large)            # ~14 times the cut cells: nothing fits any more
    prob.step()
    with pytest.raises(CfxError):
        ctx.check()
    # recovery: the same step eagerly grows the buffers and gives the right answer
    ctx.set_deferred(False)
    prob.step()
    _, _, ref = make(shape, p0, p1, large, 2)
    ref.step(keep=True)
    assert_same(results(prob), results(ref))
    # and deferred mode works again with the new capacities
    ctx.set_deferred(True)
    prob.step()
    ctx.check()
    assert_same(results(prob), results(ref))
