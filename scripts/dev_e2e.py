"""Host-side breakdown of one bt_solve_windows call on the bench workload (developer aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth
from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
dev = torch.device("cuda:0")
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 32
scenes = [synth.make_window(s) for s in range(2)]
wins = []
for k in range(nw):
    sc = scenes[k % 2]
    depth = [torch.from_numpy(sc.depth[f]).to(dev) for f in range(sc.n_frames)]
    normal = [torch.from_numpy(sc.normal[f]).to(dev) for f in range(sc.n_frames)]
    wins.append(SolveWindow(sc.corr, sc.H, sc.W, depth, normal, sc.poses_init, sc.K))
opt = OptimizerGpu(None, max_windows=nw, max_frames=10, max_corr=2000)
for _ in range(5): opt.optimizeWindows(wins)
torch.cuda.synchronize()
R = 200
def T(fn):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(R): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / R * 1e3
print("optimizeWindows  ms", T(lambda: opt.optimizeWindows(wins)))
print("marshal only     ms", T(lambda: opt._marshal(wins)))
def st():
    opt.stage(wins); torch.cuda.synchronize()
print("stage+sync       ms", T(st))
def st2():
    opt.stage(wins)
print("stage (async)    ms", T(st2))
def rn():
    opt.run(); torch.cuda.synchronize()
print("run+sync         ms", T(rn))
print("fetch            ms", T(lambda: opt.fetch()))
opt.close()
