"""Developer diagnostic (run under gpurun): CUDA solver vs oracle A (CPU) vs oracle B (reference kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bundletrack_b200 import synth
from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
import oracle

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
for seed, N, C in ((0, 10, 2000), (1, 2, 500), (2, 5, 800)):
    w = synth.make_window(seed, n_frames=N, n_corr=C)
    depth = [torch.from_numpy(w.depth[k]).to(dev) for k in range(N)]
    normal = [torch.from_numpy(w.normal[k]).to(dev) for k in range(N)]
    opt = OptimizerGpu(None, max_windows=4, max_frames=15, max_corr=8192)
    opt.enable_debug(True)
    win = SolveWindow(w.corr, w.H, w.W, depth, normal, w.poses_init, w.K)
    t = time.time(); out = opt.optimizeWindows([win])[0]; torch.cuda.synchronize(); t1 = time.time() - t
    t = time.time(); out = opt.optimizeWindows([win])[0]; torch.cuda.synchronize(); t2 = time.time() - t
    print(f"seed {seed} N={N} C={C}: first call {t1*1e3:.2f} ms, second {t2*1e3:.3f} ms, stats {opt.stats()}")
    oa = oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init)
    print("  cuda vs oracleA f32:", synth.pose_errors(out, oa))
    oa64 = oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, precision="f64")
    print("  cuda vs oracleA f64:", synth.pose_errors(out, oa64))
    # dense system of the last GN iteration vs oracle at the poses entering that iteration
    p6 = oracle.default_params(num_iter_outer=6)
    o6 = oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, params=p6)
    JtJ, Jtr, nf = oracle.dense_system(w.depth, w.normal, w.K, o6)
    J, r = opt.debug_dense(0, N)
    print("  dense JtJ rel err %.3e, Jtr rel err %.3e (|JtJ| %.3e)" % (np.abs(J - JtJ).max() / np.abs(JtJ).max(), np.abs(r - Jtr).max() / max(np.abs(Jtr).max(), 1e-30), np.abs(JtJ).max()))
    try:
        ob, pairs, t_outer, t_solve = oracle.ref_optimize_frames([d.data_ptr() for d in depth], [n.data_ptr() for n in normal], w.H, w.W, w.K, w.corr, w.poses_init)
        ob2, pairs2, t_outer2, t_solve2 = oracle.ref_optimize_frames([d.data_ptr() for d in depth], [n.data_ptr() for n in normal], w.H, w.W, w.K, w.corr, w.poses_init)
        print("  oracleB pairs (tgt,src) first 6:", pairs[:6].tolist(), " all tgt>src:", bool((pairs[:, 0] > pairs[:, 1]).all()), f" t_outer {t_outer2:.2f} ms t_solve {t_solve2:.2f} ms")
        print("  oracleB run-to-run:", synth.pose_errors(ob, ob2))
        print("  oracleA(pairs from B) vs oracleB:", synth.pose_errors(oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, pairs=pairs), ob))
        win_b = SolveWindow(w.corr, w.H, w.W, depth, normal, w.poses_init, w.K, dense_pairs=pairs)
        out_b = opt.optimizeWindows([win_b])[0]
        print("  cuda(pairs from B) vs oracleB:", synth.pose_errors(out_b, ob))
    except Exception as e:
        print("  oracle B failed:", repr(e))
    print("  err vs GT: init", synth.pose_errors(w.poses_init, w.poses_gt), "cuda", synth.pose_errors(out, w.poses_gt))
    opt.close()
