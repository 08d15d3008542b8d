"""Generates tests/golden/ref_window_*.npz ON A GPU BOX (run under gpurun): small synthetic windows pushed through the
REFERENCE'S OWN CUDA kernels (oracle/_ref/libbt_ref.so = /root/reference/src/cuda/{Solver/SolverBundling,SBA,
CUDAImageUtil}.cu compiled verbatim) — inputs, the dense pair directions the reference picked and its
output poses.  The CPU test suite then pins oracle/solver_oracle.c against these vectors without a GPU.

    gpurun -- 'python scripts/make_golden_ref.py gpurun_out/golden'   &&  cp gpurun_out/golden/*.npz tests/golden/
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from bundletrack_b200 import synth

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
os.makedirs(out_dir, exist_ok=True)
dev = torch.device("cuda:0")
cases = [  # (seed, n_frames, n_corr, H, W, outer, inner)
    (101, 3, 300, 240, 320, 7, 5),
    (102, 4, 240, 120, 160, 7, 5),
    (103, 2, 500, 120, 160, 1, 5),   # BASELINE config 0 shape: 2 keyframes, 500 correspondences, 1 GN iteration
]
for idx, (seed, N, C, H, W, outer, inner) in enumerate(cases):
    K = tuple(v * W / 640.0 for v in synth.NOCS_K)
    w = synth.make_window(seed, n_frames=N, n_corr=C, H=H, W=W, K=K)
    depth = [torch.from_numpy(w.depth[k]).to(dev) for k in range(N)]
    normal = [torch.from_numpy(w.normal[k]).to(dev) for k in range(N)]
    prm = oracle.default_params(num_iter_outer=outer, num_iter_inner=inner)
    runs = [oracle.ref_optimize_frames([d.data_ptr() for d in depth], [n.data_ptr() for n in normal], H, W, w.K, w.corr, w.poses_init, prm) for _ in range(3)]
    poses_ref, pairs = runs[0][0], runs[0][1]
    jitter = max(max(synth.pose_errors(r[0], poses_ref)) for r in runs[1:])
    a = oracle.solve_window(w.depth, w.normal, w.K, w.corr, w.poses_init, pairs=pairs, params=prm)
    print(f"case {idx}: N={N} C={C} {W}x{H} pairs={pairs.tolist()} ref run-to-run {jitter:.2e}  oracleA vs ref {synth.pose_errors(a, poses_ref)}")
    np.savez_compressed(os.path.join(out_dir, f"ref_window_{idx}.npz"), depth=w.depth, normal=w.normal, K=np.asarray(w.K, np.float32),
                        corr=w.corr.view(np.uint8).reshape(-1, 32), poses_init=w.poses_init, poses_ref=poses_ref, pairs=pairs,
                        num_iter_outer=outer, num_iter_inner=inner, ref_jitter=jitter)
