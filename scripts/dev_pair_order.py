"""Does the ORDER of the dense pair list change the result?  (developer aid: the oracle is insensitive to it)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import matcher_oracle as mo
from bundletrack_b200 import synth
from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
dev = torch.device("cuda:0")
u3 = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "curand_xorwow_seed0.npy"))[:2000]
prm = (0.02, np.cos(np.deg2rad(45)), 10000, np.cos(np.deg2rad(180)))
seed = 5
w = synth.make_window(seed, n_frames=4, n_corr=10)
fr = synth.make_feature_frames(w, 500, seed=seed)
host = [{"kpts": fr[k]["kpts"], "desc": fr[k]["desc"], "depth": w.depth[k], "normal": w.normal[k], "pose": w.poses_init[k], "id": k} for k in range(4)]
ents = []
for i in range(4):
    for j in range(i + 1, 4):
        rows, ids = mo.find_corres(host[j], host[i], w.K, prm, u3, 0.01)
        if ids is not None:
            for r in rows[ids]: ents.append((i, j, r[7:10], r[4:7]))
corr = np.zeros(len(ents), synth.ENTRYJ_DTYPE)
for k, e in enumerate(ents): corr[k] = e
depth = [torch.from_numpy(w.depth[k]).to(dev) for k in range(4)]
normal = [torch.from_numpy(w.normal[k]).to(dev) for k in range(4)]
o = OptimizerGpu(None, max_windows=1, max_frames=8, max_corr=8192)
o.enable_debug(True)
import itertools
base = [[1, 0], [3, 1], [2, 1], [2, 0], [3, 0], [3, 2]]
for perm in ([0, 1, 2, 3, 4, 5], [0, 2, 1, 3, 4, 5], [5, 4, 3, 2, 1, 0], [1, 0, 2, 3, 4, 5], [0, 1, 2, 3, 5, 4]):
    pairs = np.array([base[i] for i in perm], np.uint32)
    out = o.optimizeWindows([SolveWindow(corr, w.H, w.W, depth, normal, w.poses_init, w.K, dense_pairs=pairs)])[0]
    cnt = o.debug_counts(0, len(pairs))
    a = oracle.solve_window(w.depth, w.normal, w.K, corr, w.poses_init, pairs=pairs)
    print(perm, "lib vs oracle %.2e" % max(synth.pose_errors(out, a)), "counts last iter", cnt.astype(int).tolist())
o.close()
