"""Developer check of the tcgen05 matcher on the GPU box: correctness vs the exact oracle, timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth
from bundletrack_b200.matcher import KnnMatcher
from oracle import matcher_oracle as mo
dev = torch.device("cuda:0")
m = KnnMatcher(max_pairs=64, max_feats=5120)
m.enable_timing(True)
for (na, nb) in ((100, 90), (500, 500), (2000, 2000), (5000, 5000), (1, 700), (300, 5)):
    a, b, ia, ib = synth.make_descriptors(na + nb, na, nb)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    iAB, dAB, iBA, dBA = m.knn_match_pairs([(ta, tb)])
    torch.cuda.synchronize()
    tm = m.timing()
    i1, d1 = mo.knn(a, b); i2, d2 = mo.knn(b, a)
    gi, gd = iAB[0].cpu().numpy(), dAB[0].cpu().numpy()
    hi, hd = iBA[0].cpu().numpy(), dBA[0].cpu().numpy()
    okA = (gi == i1).all(); okB = (hi == i2).all()
    fin = np.isfinite(d1)
    print(f"{na}x{nb}: AB idx match {okA} ({(gi==i1).mean():.5f}) maxd {np.abs(gd[fin]-d1[fin]).max() if fin.any() else 0:.2e} | BA idx match {okB} ({(hi==i2).mean():.5f}) | {tm}")
    if not okA:
        bad = np.nonzero((gi != i1).any(1))[0][:5]
        for r in bad: print("   row", r, "got", gi[r], gd[r], "want", i1[r], d1[r])
    # timing over repeats
    for _ in range(3): m.knn_match_pairs([(ta, tb)])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): m.knn_match_pairs([(ta, tb)])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    tm = m.timing()
    fl = 2.0 * na * nb * 256 * 2
    print(f"   wall {dt*1e6:.1f} us/call; tc kernel {tm['tc_ms']*1e3:.1f} us -> {fl/ (tm['tc_ms']*1e-3) / 1e12:.1f} TFLOP/s (both directions executed)")
# batch of 45 pairs among 10 frames, 2000 feats (cfg2)
frames = [torch.from_numpy(synth.make_descriptors(100 + f, 2000, 8)[0]).to(dev) for f in range(10)]
pairs = [(frames[j], frames[i]) for i in range(10) for j in range(i + 1, 10)]
for _ in range(3): out = m.knn_match_pairs(pairs)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): out = m.knn_match_pairs(pairs)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
tm = m.timing()
print(f"cfg2 45 pairs x 2000: wall {dt*1e3:.3f} ms/call; {tm}; tc {2*45*2.0*2000*2000*256/(tm['tc_ms']*1e-3)/1e12:.1f} TFLOP/s")
i1, d1 = mo.knn(frames[1].cpu().numpy(), frames[0].cpu().numpy())
print("  pair0 AB match:", (out[0][0].cpu().numpy() == i1).all())
