"""Where the end-to-end time of bench.py's C2 line goes: H2D of u0, the steps, the streaming save (device snapshot +
D2H), each timed on the host with a synchronize in between; plus the raw pinned-memory PCIe rates of this box."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import ggp_b200 as G
import problems as P

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
x = torch.empty(N * N * 2, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); h2d = time.perf_counter() - t
t = time.perf_counter(); x.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h = time.perf_counter() - t
print(f"torch pinned {x.numel()*4/1e6:.1f} MB: H2D {h2d*1e3:.3f} ms ({x.numel()*4/h2d/1e9:.1f} GB/s)  D2H {d2h*1e3:.3f} ms ({x.numel()*4/d2h/1e9:.1f} GB/s)")

pb = P.kerr2d(G, N=N, nsteps=steps)
prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
it = G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, show_progress=False)
lib = it.lib
it.upload(); it.advance(5); it.save_async(0); it.save_wait()
for rep in range(3):
    t0 = time.perf_counter(); it.upload(); t1 = time.perf_counter()
    it.advance(steps); t2 = time.perf_counter()
    lib.ggp_synchronize(it.handle); t3 = time.perf_counter()
    it.save_async(0); t4 = time.perf_counter()
    it.save_wait(); t5 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.3f}  advance(issue) {1e3*(t2-t1):.3f}  steps(sync) {1e3*(t3-t2):.3f}  save_async(issue) {1e3*(t4-t3):.3f}  save_wait {1e3*(t5-t4):.3f}  total {1e3*(t5-t0):.3f} ms")
for rep in range(3):
    t0 = time.perf_counter(); it.upload(); it.advance(steps); it.save_async(0); it.save_wait(); t5 = time.perf_counter()
    print(f"pipelined rep {rep}: total {1e3*(t5-t0):.3f} ms  -> {N*N*steps/(t5-t0)/1e9:.2f} G pt-steps/s")
