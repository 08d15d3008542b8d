"""Critical path of ONE window through k_solve on the device-wide timer (developer aid; needs the variant library built with
-DBT_PROF_GLOBALTIMER: BT_B200_LIB=bundletrack_b200/lib/variants/libbt_gtimer.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth
from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
dev = torch.device("cuda:0")
sc = synth.make_window(0)
depth = [torch.from_numpy(sc.depth[f]).to(dev) for f in range(sc.n_frames)]
normal = [torch.from_numpy(sc.normal[f]).to(dev) for f in range(sc.n_frames)]
opt = OptimizerGpu(None, max_windows=1, max_frames=10, max_corr=2000)
opt.stage([SolveWindow(sc.corr, sc.H, sc.W, depth, normal, sc.poses_init, sc.K)])
for _ in range(3): opt.run()
torch.cuda.synchronize()
opt.enable_profile(100000)
opt.run(); torch.cuda.synchronize()
rec = opt.get_profile()
kind = rec[:, 0] >> 32
tiles, tails = rec[kind == 0], rec[kind == 1]
it_of_tile = (tiles[:, 1] >> 16) & 0xffff
idx_of_tile = tiles[:, 1] >> 32
it_of_tail = tails[:, 1] >> 32
t00 = tiles[:, 2].min()
print("iteration: first tile work start | last dense tile end | sparse tile: start..end | tail: start .. P0/P2 .. gathers .. diag .. cross .. PCG .. update end   (us from kernel start)")
for it in range(7):
    tl = tiles[it_of_tile == it]
    sp = tl[idx_of_tile[it_of_tile == it] < 2]      # the window's two sparse tiles come first
    dn = tl[idx_of_tile[it_of_tile == it] >= 2]
    tt = tails[it_of_tail == it][0]
    f = lambda v: f"{(v - t00) / 1e3:7.2f}"
    print(it, f(dn[:, 3].min()), "|", f(dn[:, 6].max()), "| sparse", f(sp[:, 3].min()), f(sp[:, 6].max()), "| tail", f(tt[2]), " ".join(f(tt[2 + k]) for k in (1, 3, 4, 5, 6, 7, 8)),
          "| dense tile dur p50/max us", f"{np.median(dn[:,6]-dn[:,3])/1e3:.2f}/{(dn[:,6]-dn[:,3]).max()/1e3:.2f}", "sparse dur", f"{(sp[:,6]-sp[:,3]).max()/1e3:.2f}")
it = 3
tl = tiles[it_of_tile == it]
cta = tl[:, 0] & 0xffffffff
st = (tl[:, 3] - tl[:, 3].min()) / 1e3; claim = (tl[:, 2] - t00) / 1e3; dep = (tl[:, 3] - t00) / 1e3; en = (tl[:, 6] - t00) / 1e3
order = np.argsort(st)
print("iteration 3: pixel-loop start relative to the earliest, percentiles 50/90/99/max:", np.percentile(st, [50, 90, 99, 100]).round(2).tolist())
print(" latest starters (tile idx, cta, claim us, start us, end us, px):")
for j in order[-12:]:
    print("  ", int(tl[j, 1] >> 32), int(cta[j]), f"{claim[j]:.2f} {dep[j]:.2f} {en[j]:.2f}", int(tl[j, 1] & 0xffff))
# which CTAs ran more than one tile of this iteration?
u, c = np.unique(cta, return_counts=True)
print(" CTAs with 2+ tiles in iteration 3:", int((c > 1).sum()), "max", int(c.max()), "; CTAs with a tile:", len(u))
opt.close()
