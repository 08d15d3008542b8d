"""Per-iteration divergence study on the GPU box: CUDA vs oracle A (f32/f64) vs oracle B, gating counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth
from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
import oracle
dev = torch.device("cuda:0")
def yml(it): return {"bundle": {"num_iter_outter": it, "num_iter_inner": 5, "robust_delta": 0.005, "image_downscale": 4}, "p2p": {"max_dist": 0.02, "max_normal_angle": 45}}
for seed, N, C in ((9, 15, 3000), (0, 10, 2000), (40, 6, 1200)):
    w = synth.make_window(seed, n_frames=N, n_corr=C)
    depth = [torch.from_numpy(w.depth[k]).to(dev) for k in range(N)]
    normal = [torch.from_numpy(w.normal[k]).to(dev) for k in range(N)]
    dp = [d.data_ptr() for d in depth]; nq = [n.data_ptr() for n in normal]
    print(f"== seed {seed} N={N} C={C}")
    pairs = oracle.default_pairs(N)
    for it in range(1, 8):
        o = OptimizerGpu(yml(it), max_windows=1, max_frames=15, max_corr=8192); o.enable_debug(True)
        out = o.optimizeWindows([SolveWindow(w.corr, w.H, w.W, depth, normal, w.poses_init, w.K)])[0]
        prm = oracle.default_params(num_iter_outer=it)
        a32 = oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, params=prm)
        a64 = oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, params=prm, precision="f64")
        # gating counts of the last iteration: oracle at the poses entering it
        prev = w.poses_init if it == 1 else oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, params=oracle.default_params(num_iter_outer=it - 1))
        _, _, nf = oracle.dense_system(w.depth, w.normal, w.K, prev, pairs=pairs)
        cnt = o.debug_counts(0, len(pairs))
        ob, pb, _, _ = oracle.ref_optimize_frames(dp, nq, w.H, w.W, w.K, w.corr, w.poses_init, prm)
        ob2, _, _, _ = oracle.ref_optimize_frames(dp, nq, w.H, w.W, w.K, w.corr, w.poses_init, prm)
        outb = o.optimizeWindows([SolveWindow(w.corr, w.H, w.W, depth, normal, w.poses_init, w.K, dense_pairs=pb)])[0]
        ab = oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, params=prm, pairs=pb)
        e = lambda x, y: "%.1e/%.1e" % synth.pose_errors(x, y)
        print(f" it{it}: cuda-A32 {e(out,a32)} A32-A64 {e(a32,a64)} | count diff sum|.|={int(np.abs(cnt-nf).sum())} of {int(nf.sum())} | B-B {e(ob,ob2)} cuda-B {e(outb,ob)} A32-B {e(ab,ob)}")
        o.close()
