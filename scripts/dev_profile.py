"""In-kernel phase profile of k_solve on the GPU box (developer aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth
from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
dev = torch.device("cuda:0")
nwin = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,32").split(",")]
scenes = [synth.make_window(s) for s in range(2)]
for nw in nwin:
    wins = []
    for k in range(nw):
        sc = scenes[k % 2]
        depth = [torch.from_numpy(sc.depth[f]).to(dev) for f in range(sc.n_frames)]
        normal = [torch.from_numpy(sc.normal[f]).to(dev) for f in range(sc.n_frames)]
        wins.append(SolveWindow(sc.corr, sc.H, sc.W, depth, normal, sc.poses_init, sc.K))
    opt = OptimizerGpu(None, max_windows=nw, max_frames=10, max_corr=2000)
    opt.enable_timing(True)
    opt.stage(wins)
    for _ in range(3): opt.run()
    torch.cuda.synchronize()
    print(f"== {nw} windows: kernel ms {opt.timing_ms()} stats {opt.stats()}")
    opt.enable_profile(400000)
    opt.run(); torch.cuda.synchronize()
    rec = opt.get_profile()
    kind = rec[:, 0] >> 32
    tiles, tails = rec[kind == 0], rec[kind == 1]
    if len(tiles):
        d = np.diff(tiles[:, 2:8], axis=1)
        names = ["wait", "Msetup", "pixels", "reduce+store", "ticket"]
        print(f" tiles: n={len(tiles)} mean px/tile {np.mean(tiles[:,1] & 0xffff):.0f}; cycles mean/p50/p95 per phase:")
        for i, nm in enumerate(names):
            print(f"   {nm:14s} {d[:, i].mean():9.0f} {np.percentile(d[:, i], 50):9.0f} {np.percentile(d[:, i], 95):9.0f}")
        print(f"   total          {(tiles[:,7]-tiles[:,2]).mean():9.0f}")
    pl = rec[kind == 3]
    if len(pl):
        print(' plan phases (cycles): counts, scan, pair prefix, tile records, tail:', np.diff(pl[0, 2:8]).tolist())
    ws = rec[kind == 2]
    if len(ws):
        px = ws[:, 1] >> 32
        print(f" ws tiles: n={len(ws)} mean px {px.mean():.0f} ready-frac {np.mean(ws[:,1] & 1):.3f}")
        for nm, v in (("pixels", ws[:, 3] - ws[:, 2]), ("phaseA total", ws[:, 4] - ws[:, 2]), ("h reduce", ws[:, 6] - ws[:, 5]), ("h claim", ws[:, 7] - ws[:, 6]),
                      ("h dep+setup", ws[:, 8] - ws[:, 7]), ("h total", ws[:, 8] - ws[:, 5])):
            print(f"   {nm:14s} {v.mean():9.0f} {np.percentile(v, 50):9.0f} {np.percentile(v, 95):9.0f}")
    d = np.diff(tails[:, 2:11], axis=1)
    names = ["P0 zero/T", "P1 sparse", "P2 pairsum", "P3a gathers", "P3b diag/rhs", "P4 cross", "PCG", "update"]
    print(f" tails: n={len(tails)}")
    for i, nm in enumerate(names):
        print(f"   {nm:14s} {d[:, i].mean():9.0f} {np.percentile(d[:, i], 50):9.0f} {np.percentile(d[:, i], 95):9.0f}")
    print(f"   total          {(tails[:,10]-tails[:,2]).mean():9.0f}")
    opt.close()
