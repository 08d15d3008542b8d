"""Dense-term-only window (test_edge_cases): library vs oracle A vs the reference's own kernels (developer aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from bundletrack_b200 import synth
from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
dev = torch.device("cuda:0")
o = OptimizerGpu(None, max_windows=1, max_frames=8, max_corr=2000)
for seed, N in ((6, 3), (7, 3), (8, 4), (9, 5), (10, 3), (11, 6)):
    w = synth.make_window(seed, n_frames=N, n_corr=60)
    c = w.corr[:0]
    depth = [torch.from_numpy(w.depth[k]).to(dev) for k in range(N)]
    normal = [torch.from_numpy(w.normal[k]).to(dev) for k in range(N)]
    dp, nq = [d.data_ptr() for d in depth], [n.data_ptr() for n in normal]
    ref, pairs, _, _ = oracle.ref_optimize_frames(dp, nq, w.H, w.W, w.K, c, w.poses_init)
    ref2 = oracle.ref_optimize_frames(dp, nq, w.H, w.W, w.K, c, w.poses_init)[0]
    out = o.optimizeWindows([SolveWindow(c, w.H, w.W, depth, normal, w.poses_init, w.K, dense_pairs=pairs)])[0]
    a32 = oracle.solve_window(w.depth, w.normal, w.K, c, w.poses_init, pairs=pairs)
    a64 = oracle.solve_window(w.depth, w.normal, w.K, c, w.poses_init, pairs=pairs, precision="f64")
    f = lambda x, y: "%.1e" % max(synth.pose_errors(x, y))
    print(f"seed {seed} N {N}: ours-ref {f(out, ref)} ours-A32 {f(out, a32)} A32-ref {f(a32, ref)} A64-ref {f(a64, ref)} A32-A64 {f(a32, a64)} ref-ref {f(ref2, ref)}")
o.close()
