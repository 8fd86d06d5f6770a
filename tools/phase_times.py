#!/usr/bin/env python
"""Per-phase wall times (synchronised) of the multi-rank step: torchrun --nproc-per-node N tools/phase_times.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cutfemx_b200 import parallel as P  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pipe = P.RankPipeline([n] * 3, [0.0] * 3, [1.0] * 3, world, rank, lr, "sphere", (0.5, 0.5, 0.5, 0.35, 0.0), order=4)
tr = P.TorchDistTransport()
P.plan([pipe], tr)
T = {}


def tick(name, t0):
    torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()


for it in range(8):
    if it == 3:
        T.clear()
    dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    a, L = pipe.prob.build_forms()
    t = tick("build_forms", t)
    from cutfemx_b200 import fem as _fem

    pipe.Ag = _fem.create_ghost_row_pattern(a, pipe.imap.n_owned, pipe.Ag)
    t = tick("ghost_row_pattern", t)
    rng = pipe.imap.owner_ranges()
    cuts = [b for _, _, b in rng[:-1]]
    res = P.coo_of_rows(pipe.Ag.indptr_device(), pipe.Ag.indices_device(), pipe.imap.n_owned, pipe.imap.n_total, cuts=cuts) if cuts else (*P.coo_of_rows(pipe.Ag.indptr_device(), pipe.Ag.indices_device(), pipe.imap.n_owned, pipe.imap.n_total), [])
    rows, cols, offs = res
    bounds = [0] + offs + [int(rows.numel())]
    sends = pipe.mx.begin(rows, cols, {q: (bounds[k], bounds[k + 1]) for k, (q, _, _) in enumerate(rng)})
    t = tick("coo+begin", t)
    recv = tr.exchange([sends])[0]
    t = tick("exchange1", t)
    xr, xc = pipe.mx.inserted_entries(recv)
    _fem.insert_pattern_entries(a, xr, xc)
    t = tick("map+insert", t)
    pipe.prob.assemble()
    t = tick("assemble", t)
    A, b = pipe.prob.A, pipe.prob.b
    pipe.mx.finish(lambda r, c: pipe.ops.positions(A, r, c))
    t = tick("positions", t)
    vals = A.values_device()
    sends = {}
    for q in sorted(set(pipe.mx.send_pos) | set(pipe.vx.send_sel)):
        parts = []
        if q in pipe.mx.send_pos:
            parts.append(pipe.ops.gather(vals, pipe.mx.send_pos[q]))
        if q in pipe.vx.send_sel:
            parts.append(b[pipe.vx.send_sel[q]])
        sends[q] = torch.cat(parts)
    t = tick("pack", t)
    recv = tr.exchange([sends], counts=[pipe.recv_counts()])[0]
    t = tick("exchange2", t)
    pipe.phase_c(recv)
    t = tick("unpack", t)
    pipe.finish_step()
    t = tick("release", t)
if rank == 0:
    tot = sum(T.values()) / 5
    print(f"n={n} world={world} total {tot:.3f} ms/step (host-synchronised phases)")
    for k, v in T.items():
        print(f"  {k:20s} {v/5:7.3f} ms")
dist.destroy_process_group()
