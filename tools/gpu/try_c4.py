"""Scratch driver: the elasticity pipeline (configs[3]) at a given n, eager then graph-captured; prints stage times."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cutfemx_b200 import parallel as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = sys.argv[2] if len(sys.argv) > 2 else "eager"
t0 = time.perf_counter()
pipe = P.RankPipeline([n] * 3, [0.0] * 3, [1.0] * 3, 1, 0, 0, "torus", (0.5, 0.5, 0.5, 0.3, 0.12, 0.0), order=4, degree=2,
                      problem="elasticity", bs=3)
prob, ctx = pipe.prob, pipe.ctx
pipe.xplan = None
prob.persistent = True
torch.cuda.synchronize()
print("setup", time.perf_counter() - t0, flush=True)
for i in range(3):
    t0 = time.perf_counter()
    pipe.step_static()
    torch.cuda.synchronize()
    print("eager step", i, (time.perf_counter() - t0) * 1e3, "ms", flush=True)
print(prob.fetch_stats(), "device GB", ctx.device_bytes / 1e9, flush=True)
ctx.stage_timing(True)
ctx.stage_reset()
pipe.step_static()
torch.cuda.synchronize()
for name, ms, by in ctx.stages():
    print(f"{name:28s} {ms:9.3f} ms  {by/1e9:8.3f} GB")
ctx.stage_timing(False)
ctx.stage_reset()
if mode == "graph":
    margin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
    prob.persistent = True
    ctx.set_deferred(False, margin)
    for _ in range(2):
        pipe.step_static()
        torch.cuda.synchronize()
        print("eager+margin device GB", ctx.device_bytes / 1e9, "torch GB", torch.cuda.memory_allocated() / 1e9, flush=True)
    ctx.set_deferred(True)
    pipe.step_static()
    torch.cuda.synchronize()
    ctx.check()
    print("deferred device GB", ctx.device_bytes / 1e9, flush=True)
    ctx.graph_begin()
    try:
        pipe.step_static()
    finally:
        prob.graph = g = ctx.graph_end()
    print("captured device GB", ctx.device_bytes / 1e9, flush=True)
    for _ in range(3):
        prob.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        prob.replay()
    e1.record()
    torch.cuda.synchronize()
    ctx.check()
    print("graph step", e0.elapsed_time(e1) / 5, "ms", "kernels", g.kernel_nodes)
