#!/usr/bin/env python
"""bench.py -- cut-cell quadrature + assembly throughput of the CutFEMx hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--n 256] [--impl ours|reference]

One "step" = one pass of the whole hot path over the synthetic workload (demo_poisson.py:156-201):
classify -> locate -> volume/interface run-time quadrature -> normals -> ghost-penalty facets ->
sparsity -> matrix + vector assembly.  Prints ONE JSON line (see DESIGN.md "Measurement").

Workloads (BASELINE.json configs / SURVEY.md section 8 sizing):
  C1  2D circle R=0.5 on [-1,1]^2, 64x64 right-diagonal triangles, P1, order 4
  C3  3D sphere R=0.35 on [0,1]^3, n^3 Kuhn tetrahedra (default n=256), P1, order 4   <- default
  C2  configs[1]: P2 u on 4096^2 triangles;  C4  configs[3]: P2-vector elasticity, torus, 192^3 tets;
  C5  configs[4]: moving sphere on 128^3 tets, everything rebuilt every step
`value` = cut cells / s with all inputs resident in HBM; `e2e` = the same through host buffers
(level-set values H2D from pinned memory, CSR pattern + values + rhs D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "C1": dict(tdim=2, n=64, p0=(-1.0, -1.0), p1=(1.0, 1.0), ls=("sphere", (0.0, 0.0, 0.0, 0.5, 0.0)), order=4,
               name="2D circle R=0.5, {n}x{n} right-diagonal triangles, P1, order 4"),
    "C2p": dict(tdim=2, n=4096, p0=(-1.0, -1.0), p1=(1.0, 1.0), ls=("sphere", (0.0, 0.0, 0.0, 0.5, 0.0)), order=4,
                name="2D circle R=0.5, {n}x{n} right-diagonal triangles, P1, order 4"),
    # configs[1]: P2 u (P1 level set) on the 4096^2 triangle mesh, ghost penalty
    "C2": dict(tdim=2, n=4096, p0=(-1.0, -1.0), p1=(1.0, 1.0), ls=("sphere", (0.0, 0.0, 0.0, 0.5, 0.0)), order=4,
               degree=2, ref_n=512,
               name="2D circle R=0.5, {n}x{n} right-diagonal triangles, P2 u / P1 level set, order 4, "
                              "Nitsche + ghost-penalty facets"),
    # configs[4]: moving sphere, everything re-cut / regenerated / reassembled every step (demo_moving_poisson.py)
    "C5": dict(tdim=3, n=128, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0), ls=("sphere", (0.3, 0.5, 0.5, 0.25, 0.0)),
               order=2, moving=(0.3, 0.4, 99),
               name="moving sphere R=0.25 (c_x = 0.3 + 0.4 t/99), {n}^3 Kuhn tetrahedra, P1, order 2: re-cut, "
                    "regenerate quadrature, rebuild sparsity and reassemble every step"),
    # configs[3]: linear elasticity on a P2 VECTOR space, torus level set, Nitsche + ghost penalty (demo_elasticity.py)
    "C4": dict(tdim=3, n=192, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0), ls=("torus", (0.5, 0.5, 0.5, 0.3, 0.12)),
               order=4, degree=2, bs=3, problem="elasticity", ref_n=32,
               name="3D torus (R=0.3, r=0.12), {n}^3 Kuhn tetrahedra, linear elasticity on a P2 vector space / P1 "
                    "level set, order 4, vector Nitsche + ghost-penalty facets"),
    "C3": dict(tdim=3, n=256, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0), ls=("sphere", (0.5, 0.5, 0.5, 0.35, 0.0)),
               order=4, name="3D sphere R=0.35, {n}^3 Kuhn tetrahedra, P1 volume+interface quadrature order 4, "
                             "Nitsche + ghost-penalty facets"),
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(kernel, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{kernel}@n{n}")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm
_SLAB = {}  # per worker process: the slab's mesh, space and level set, built once and reused by every step


def _oracle_slab_init(wl, n, nparts, counter):
    """Pool initializer: worker k builds z-slab (3D) / y-strip (2D) k of the workload -- the role of one MPI rank of
    the reference.  Mesh construction stands in for reading a DOLFINx mesh and is not part of the timed path."""
    sys.path.insert(0, ROOT)
    from cutfemx_b200 import mesh as M

    with counter.get_lock():
        rank = counter.value
        counter.value += 1
    tdim = wl["tdim"]
    lo = (n * rank) // nparts
    hi = (n * (rank + 1)) // nparts
    p0, p1 = list(wl["p0"]), list(wl["p1"])
    ax = tdim - 1
    h = (p1[ax] - p0[ax]) / n
    q0, q1 = list(p0), list(p1)
    q0[ax], q1[ax] = p0[ax] + lo * h, p0[ax] + hi * h
    mesh = M.create_box(n, n, hi - lo, q0, q1) if tdim == 3 else M.create_rectangle(n, hi - lo, q0, q1)
    Vphi = M.functionspace(mesh, 1)
    bs = wl.get("bs", 1)
    V = Vphi if (wl.get("degree", 1) == 1 and bs == 1) else M.functionspace(mesh, wl.get("degree", 1), bs=bs)
    _SLAB.update(wl=wl, mesh=mesh, Vphi=Vphi, V=V, rank=rank)


def _oracle_slab_step(tstep):
    """One pass of demo_poisson.py:156-201 through the oracle on this worker's slab."""
    from cutfemx_b200 import mesh as M
    from oracle import pipeline

    wl, mesh, Vphi, V = _SLAB["wl"], _SLAB["mesh"], _SLAB["Vphi"], _SLAB["V"]
    kind, prm = wl["ls"]
    prm = list(prm)
    if "moving" in wl:
        x0, dx, period = wl["moving"]
        prm[0] = x0 + dx * (tstep % (period + 1)) / period
    ls = M.sphere_level_set(prm[:3], prm[3]) if kind == "sphere" else M.torus_level_set(prm[:3], prm[3], prm[4])
    phi = M.interpolate(Vphi, ls)
    if wl.get("problem") == "elasticity":
        out = pipeline.run_elasticity_pipeline(mesh, Vphi.dofmap, phi, V, order=wl["order"])
    else:
        out = pipeline.run_pipeline(mesh, Vphi.dofmap, phi, V, order=wl["order"])
    return dict(time=out["total_s"], cut=int(out["cut"].size), cells=int(mesh.num_cells), nnz=int(out["cols"].size))


class CpuReference:
    """The reference's CPU path restated (oracle), one serial process per partition like its MPI model; the time of
    a step is the slowest partition's.  The partitions are ghost-free sub-boxes: the CPU arm does no exchange, which
    flatters it slightly."""

    def __init__(self, wl, n, nparts):
        import multiprocessing as mp

        import oracle

        oracle.build()
        ctx = mp.get_context("fork")
        self.nparts = nparts
        self.pool = ctx.Pool(nparts, initializer=_oracle_slab_init, initargs=(wl, n, nparts, ctx.Value("i", 0)))
        self.k = 0

    def step(self):
        res = self.pool.map(_oracle_slab_step, [self.k] * self.nparts, chunksize=1)
        self.k += 1
        return dict(time=max(r["time"] for r in res), cut=sum(r["cut"] for r in res),
                    cells=sum(r["cells"] for r in res), nnz=sum(r["nnz"] for r in res))

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_parts(n, n_gpus):
    """Processes of the CPU arm: every host core (the reference is MPI-parallel), at least 4 cell layers each."""
    return max(1, min(os.cpu_count() or 1, n // 4, 64))


def run_reference(args, wl):
    # the SAME configuration as our arm unless --ref-n (or the workload, where the serial oracle would need many
    # minutes per step) asks for a bounded sample
    n = args.ref_n or wl.get("ref_n") or args.n or wl["n"]
    n_ours = args.n or wl["n"]
    nparts = reference_parts(n, args.gpus)
    ref = CpuReference(wl, n, nparts)
    # a CPU loop has no clocks or caches worth warming for long: one untimed step at most
    warm = min(args.warmup, 1)
    times, res = [], None
    for i in range(warm + args.steps):
        res = ref.step()
        if i >= warm:
            times.append(res["time"])
    ref.close()
    t = float(np.mean(times))
    value = res["cut"] / t
    sample = (f"{'the whole workload' if n == n_ours else 'same workload at n=%d' % n} ({res['cells']} cells, "
              f"{res['cut']} cut cells) as {nparts} serial slab processes (one per host core, no ghost exchange), "
              f"{warm} warm-up + {args.steps} timed steps; CPU restatement of reference loops -- reference binary "
              f"unavailable")
    line = {
        "impl": "reference", "metric": "cut_cells_per_s", "value": value, "unit": "cut-cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"].format(n=n), "cells": res["cells"], "cut_cells": res["cut"]},
        "same_config": n == n_ours,
        "nnz_per_s": res["nnz"] / t, "total_cells_per_s": res["cells"] / t,
        "cpu_baseline": {"value": value, "unit": "cut-cells/s", "cores": nparts, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cut-cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- our arm
def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads (and so its page-locked buffers: first touch) to the NUMA node of its GPU, so the
    end-to-end leg's PCIe copies do not cross the socket interconnect.  Returns what it found (None: no NUMA
    information in this machine / VM, nothing changed)."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"pci": bus, "numa_node": None}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": bus, "numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # no sysfs entry, no permission: leave the affinity alone
        return {"numa_node": None, "why": type(e).__name__}


def V_bs(pipe):
    return getattr(pipe.V, "bs", 1) or 1


def run_ours(args, wl):
    """The whole step of every rank is ONE CUDA-graph launch: deferred sizes (no host round trip inside the step)
    and, for N > 1, the ghost exchange over NCCL inside the same graph (static exchange plan, cfx_xplan_*)."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    host_binding = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from cutfemx_b200 import fem as _fem
    from cutfemx_b200 import mesh as M
    from cutfemx_b200 import parallel as P
    from cutfemx_b200._lib import HOST, check, lib

    n = args.n or wl["n"]
    tdim = wl["tdim"]
    kind, prm = wl["ls"]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    # ---- one-time setup (mesh-only data: geometry cache, facet-cell table, incidence, static pattern tables, and
    # for N > 1 the static exchange plan); reported next to ms_per_step, not inside it
    sync()
    t0 = time.perf_counter()
    # strong scaling: the fixed n^tdim mesh is split into z-slabs (y-strips in 2D), one rank per GPU, with a
    # shared-facet ghost layer; work-balanced cuts (the role of vertex weights in DOLFINx's graph partitioner):
    # cut and inside cells cost ~40x an outside cell, and they cluster around the level set
    ls_fn = M.sphere_level_set(prm[:3], prm[3]) if kind == "sphere" else M.torus_level_set(prm[:3], prm[3], prm[4])
    ranges = P.slab_ranges(n, world, P.layer_weights([n] * tdim, wl["p0"], wl["p1"], ls_fn)) if world > 1 else None
    pipe = P.RankPipeline([n] * tdim, list(wl["p0"]), list(wl["p1"]), world, rank, local_rank, kind, prm,
                          order=wl["order"], ranges=ranges, degree=wl.get("degree", 1),
                          **({"problem": wl["problem"], "bs": wl.get("bs", 1), "p2_numbering": "blocks"}
                             if "problem" in wl else {}))
    bsz = int(V_bs(pipe))
    prob, ctx, V = pipe.prob, pipe.ctx, pipe.V
    sync()
    t1 = time.perf_counter()
    if world > 1:
        P.plan([pipe], P.TorchDistTransport(), static=True)
        P.init_nccl(ctx, rank, world)
    else:
        pipe.xplan = None
    prob.persistent = True
    pipe.step_static()  # first step: binds the space and the topology (static tables), allocates every buffer
    sync()
    t2 = time.perf_counter()
    setup = {"ms": (t2 - t0) * 1e3, "mesh_generation_and_bind_ms": (t1 - t0) * 1e3,
             "plan_and_first_step_ms": (t2 - t1) * 1e3,
             "what": "synthetic mesh generation on the device, cfx_mesh_bind (geometry cache), static exchange plan "
                     "(N > 1), and the first step (cfx_topology_bind, cfx_space_bind: incidence, static pattern and "
                     "contribution tables)"}

    tstep = [0]

    def move():
        if "moving" in wl:  # the level set moves: new nodal values (device-resident), then the whole path
            x0, dx, period = wl["moving"]
            t = tstep[0] % (period + 1)
            tstep[0] += 1
            pipe.move_level_set((x0 + dx * t / period,) + tuple(prm[1:]))

    hnd0 = ctx.handle
    lanes_on = not os.environ.get("CFX_NO_LANES")
    graph = pipe.capture_static(margin=0.25)
    setup["device_bytes"] = ctx.device_bytes
    for _ in range(max(args.warmup, 3)):
        move()
        prob.replay()
    ctx.check()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)  # 256 MiB > 126 MB L2
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    sampler.start()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (untimed)
        sync()
        ev[i][0].record()
        move()
        prob.replay()
        ev[i][1].record()
    sync()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    ctx.check()  # a capacity overflow in any replay would surface here
    # per step: the slowest rank; then summed over the steps
    t_total = sum(allmax([a.elapsed_time(b) for a, b in ev])) / 1e3
    stats = prob.fetch_stats()
    no = pipe.imap.n_owned
    nnz_owned = int(prob.A.indptr_device()[no].item())
    # global checksums of the assembled system (owned rows): equal for every N, so the scaling run verifies the
    # NCCL path against the N = 1 line
    vals_owned = prob.A.values_device()[: nnz_owned * bsz * bsz]
    cut_total, nnz_total, cells_total, active_total, launches_total, a_fro2, b_sum, b_abs = allsum(
        [stats["cut"], nnz_owned, stats["inside"] + stats["cut"] + stats["outside"],
         stats["inside"] + stats["volume_rules"], launches, float((vals_owned * vals_owned).sum()),
         float(prob.b[: no * bsz].sum()), float(prob.b[: no * bsz].abs().sum())])
    checks = {"nnz": int(nnz_total), "A_frobenius": a_fro2 ** 0.5, "b_sum": b_sum, "b_abs_sum": b_abs,
              "cut_cells": int(cut_total)}

    # ---- the same K steps with per-stage CUDA events recorded INSIDE the graph (a second capture with stage
    # timing on): per-kernel durations for the roofline block, measured live on the stream the kernels run on
    # (lanes off for this capture: concurrent kernels would stretch each other's intervals; the timed steps above ran
    # the real thing, with the independent branches of the step on lanes of their own)
    ctx.set_lanes(False)
    ctx.stage_timing(True)
    ctx.stage_reset()
    ctx.graph_begin()
    try:
        pipe.step_static()
    finally:
        gt = ctx.graph_end()
    agg = {}
    for i in range(args.steps):
        flush.zero_()
        move()
        sync()
        gt.launch()
        torch.cuda.synchronize()
        for name, msv, by in ctx.stages():
            a = agg.setdefault(name, [0.0, 0.0, 0])
            a[0] += msv
            a[1] += by
            a[2] += 1
    gt.free()
    ctx.stage_timing(False)
    ctx.stage_reset()
    # algorithmic bytes come from the exact sizes: one eager step with stage accounting (not timed)
    ctx.set_deferred(False)
    ctx.stage_timing(True)
    move()
    pipe.step_static()
    sync()
    exact_bytes, eager_ms = {}, {}
    for name, msv, by in ctx.stages():
        exact_bytes[name] = exact_bytes.get(name, 0.0) + by
        eager_ms[name] = eager_ms.get(name, 0.0) + msv
    cnt = (C.c_int64 * 4)()
    lib().cfx_space_counters(hnd0, ctx.space_index(V), cnt)
    space_counters = {"seen_slow_rows": cnt[0], "seen_noclist_rows": cnt[1], "cap_act_rows": cnt[2], "cap_band": cnt[3]}
    ctx.stage_timing(False)
    ctx.stage_reset()
    ctx.set_deferred(True)
    ctx.set_lanes(lanes_on)
    per_stage = {}
    for k, v in agg.items():
        msk = v[0] / args.steps
        by = exact_bytes.get(k, 0.0)
        per_stage[k] = {"ms_per_step": msk, "alg_GB_per_step": by / 1e9,
                        "GBps": (by / 1e9) / (msk / 1e3) if msk > 0 else None}
    stage_sum = sum(v["ms_per_step"] for k, v in per_stage.items() if not k.endswith("_kernel"))
    # load balance: every rank's own stage sum and the three stages that scale with its share of the mesh
    mine = torch.tensor([stage_sum, per_stage.get("classify", {}).get("ms_per_step", 0.0),
                         per_stage.get("gather_matrix", {}).get("ms_per_step", 0.0),
                         per_stage.get("create_sparsity", {}).get("ms_per_step", 0.0)], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(v), 4) for v in t.tolist()] for t in per_rank]

    # ---- end-to-end leg: host buffers in, host buffers out, through the public API
    vals = pipe.phi.x.array
    h_phi = torch.empty(vals.shape, dtype=torch.float64, pin_memory=True)
    h_phi.copy_(vals)
    torch.cuda.synchronize()
    hnd = ctx.handle
    n_phi = int(vals.shape[0])  # the level set lives on the P1 vertex space
    # re-bind level set 0 to the pinned host array: cfx_update (inside the graph) now does the H2D copy every step
    check(hnd, lib().cfx_levelset_bind(hnd, 0, None, tdim + 1, 1, C.c_void_p(h_phi.data_ptr()), C.c_int64(n_phi), HOST,
                                       1))
    # Results leave through a copy stream into pinned host buffers, double-buffered: the device->host copy of step k
    # (CSR pattern + values + rhs) overlaps the compute of step k+1; every step's inputs still arrive from the host
    # and every step's results still reach it inside the timed region.  One graph per buffer set.
    nnz_cap = int(stats["nnz"] * 1.3) + 4096
    e2e_bytes = (4 + 8 * bsz * bsz) * nnz_cap + 16 * V.num_dofs * bsz
    skip_e2e = e2e_bytes > args.e2e_cap_gb * 1e9  # two pinned host buffer sets of this size would be needed
    t_e2e = float('nan')
    d2h = [0]
    if not skip_e2e:
        hb = [dict(vals=torch.empty(nnz_cap * bsz * bsz, dtype=torch.float64, pin_memory=True),
                   cols=torch.empty(nnz_cap, dtype=torch.int32, pin_memory=True),
                   rp=torch.empty(V.num_dofs + 1, dtype=torch.int64, pin_memory=True),
                   b=torch.empty(V.num_dofs * bsz, dtype=torch.float64, pin_memory=True)) for _ in range(2)]
        As = [prob.A, _fem.MatrixCSR(ctx)]
        bs = [prob.b, torch.empty_like(prob.b)]
        graphs = []
        for k in range(2):
            prob.A, prob.b = As[k], bs[k]
            graphs.append(pipe.capture_static(margin=0.25))
        side = torch.cuda.Stream()
        done = [None, None]
        kstep = [0]
        d2h = [0]

        def e2e_step():
            k = kstep[0] % 2
            kstep[0] += 1
            if done[k] is not None:
                torch.cuda.current_stream().wait_event(done[k])  # buffers of step k-2 have left the device
            if "moving" in wl:  # the host owns the level set in this leg: new values are written into the pinned array
                move()
                h_phi.copy_(pipe.phi.x.array)
            graphs[k].launch()
            A = As[k]
            A._cache.clear()
            ready = torch.cuda.Event()
            ready.record()
            nnz = A.nnz  # the host needs the size of what it receives: the one round trip of the step
            side.wait_event(ready)
            A.copy_to_host_async(hb[k]["rp"], hb[k]["cols"], hb[k]["vals"], side)
            with torch.cuda.stream(side):
                hb[k]["b"].copy_(bs[k], non_blocking=True)
                done[k] = torch.cuda.Event()
                done[k].record()
            d2h[0] = (4 + 8 * bsz * bsz) * nnz + 8 * (V.num_dofs + 1) + 8 * V.num_dofs * bsz
            return nnz

        for _ in range(2):
            e2e_step()
        side.synchronize()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        for ev_done in done:
            torch.cuda.current_stream().wait_event(ev_done)
        e1.record()
        sync()
        side.synchronize()
        ctx.check()
        t_e2e = allmax([e0.elapsed_time(e1) / 1e3])[0]
    h2d, d2h_total = allsum([8.0 * n_phi, float(d2h[0])])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    # the roofline object describes ONE kernel: among the stages that time a single kernel (named *_kernel; the
    # other stages are sequences of kernels) take the one with the largest live time
    cand = {k: v for k, v in per_stage.items() if k.endswith("_kernel")} or per_stage
    dom = max(cand.items(), key=lambda kv: kv[1]["ms_per_step"])
    roof = {"bound": "hbm", "kernel": dom[0], "achieved": dom[1]["GBps"], "peak": peak, "unit": "GB/s",
            "frac": (dom[1]["GBps"] or 0.0) / peak,
            # dram bytes of one launch from the committed ncu --set full capture: N = 1 only (no per-N capture)
            "traffic": load_traffic(dom[0], n) if world == 1 else None, "peak_source": peak_src,
            "ms_per_launch": dom[1]["ms_per_step"], "alg_GB_per_launch": dom[1]["alg_GB_per_step"],
            "rank": 0}
    line = {
        "metric": "cut_cells_per_s", "value": cut_total * args.steps / t_total, "unit": "cut-cells/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"].format(n=n), "cells": int(cells_total), "cut_cells": int(cut_total),
                   "active_cells": int(active_total), "nnz": int(nnz_total), "l2": "flushed between timed steps "
                   "(256 MiB write); inputs 2.2 GB >> L2",
                   "partition": f"{world} work-balanced slabs along the last axis {pipe.ranges}, shared-facet ghost "
                                f"layer, ghost rows exchanged over NCCL inside the step's graph (fixed-size messages)"
                   if world > 1 else "1 rank (no exchange)",
                   "step": "one CUDA-graph launch per rank and step (deferred sizes, no host round trip inside it); "
                           + ("independent branches of the step (volume rules | interface rules + normals | "
                              "ghost-penalty facets | cell lists; static | band | inactive rows of the pattern and "
                              "of the gather) are parallel branches of the graph; `stages` are timed in a second, "
                              "serial capture, so their sum exceeds ms_per_step" if lanes_on else "serial (CFX_NO_LANES)")},
        "nnz_per_s": nnz_total * args.steps / t_total, "total_cells_per_s": cells_total * args.steps / t_total,
        "e2e": ({"value": cut_total * args.steps / t_e2e, "unit": "cut-cells/s", "h2d_bytes_per_step": int(h2d),
                 "d2h_bytes_per_step": int(d2h_total), "ms_per_step": t_e2e / args.steps * 1e3} if not skip_e2e else
                {"value": None, "unit": "cut-cells/s", "skipped": f"the CSR of one step is {e2e_bytes / 1e9:.1f} GB: two "
                 f"pinned host buffer sets exceed --e2e-cap-gb {args.e2e_cap_gb}"}),
        "gpu_launches": int(launches_total), "graph_launches_per_step": 1, "kernels_per_step": graph.kernel_nodes,
        "clocks": clocks, "roofline": roof, "stages": per_stage, "stage_sum_ms": stage_sum, "setup": setup,
        "checks": checks, "space_counters": space_counters, "host_binding": host_binding,
        "per_rank_ms": {"columns": ["stage_sum", "classify", "gather_matrix", "create_sparsity"], "ranks": per_rank},
        "stages_one_eager_step_ms": {k: round(v, 4) for k, v in eager_ms.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        # one step of the reference arm on the SAME configuration (bench.py --impl reference times K of them)
        nb = args.ref_n or wl.get("ref_n") or n
        nparts = reference_parts(nb, 1)
        ref = CpuReference(wl, nb, nparts)
        r = ref.step()
        ref.close()
        line["cpu_baseline"] = {
            "value": r["cut"] / r["time"], "unit": "cut-cells/s", "cores": nparts, "kind": "port",
            "sample": f"{'the whole workload' if nb == n else 'same workload at n=%d' % nb} ({r['cells']} cells, "
                      f"{r['cut']} cut cells), one step as {nparts} serial slab processes; {r['time']:.2f} s; CPU "
                      f"restatement of reference loops -- reference binary unavailable",
            "total_cells_per_s": r["cells"] / r["time"], "nnz_per_s": r["nnz"] / r["time"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=0,
                    help="override the mesh resolution of the workload (--mesh-n under torchrun, whose own parser "
                         "takes --n for an abbreviation of its options)")
    ap.add_argument("--e2e-cap-gb", type=float, default=8.0,
                    help="skip the host-buffer leg when one step's CSR is larger than this (pinned host memory)")
    ap.add_argument("--ref-n", type=int, default=0, help="resolution of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
