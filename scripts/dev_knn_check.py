"""Developer check of the kNN matcher on the GPU box: exactness against the numpy oracle on a ladder of shapes (small first, so a
broken kernel shows up before the big launches), then device timings of cfg5 (5000 x 5000) and cfg2 (45 pairs x 2000^2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bundletrack_b200 import synth
from bundletrack_b200.matcher import KnnMatcher
from oracle import matcher_oracle as mo
dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
m = KnnMatcher(max_pairs=48, max_feats=5120)
m.enable_timing(True)
bad = 0
for na, nb in ((64, 64), (128, 256), (129, 257), (500, 500), (5, 3), (1, 700), (300, 2), (1000, 3000), (2000, 2000)):
    a, b, _, _ = synth.make_descriptors(na * 7 + nb, na, nb)
    iAB, dAB, iBA, dBA = m.knn_match_pairs([(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))])
    torch.cuda.synchronize()
    i1, d1 = mo.knn(a, b); i2, d2 = mo.knn(b, a)
    okAB = np.array_equal(iAB[0].cpu().numpy(), i1); okBA = np.array_equal(iBA[0].cpu().numpy(), i2)
    tm = m.timing()
    print(f"{na:5d} x {nb:5d}: A->B {'ok' if okAB else 'MISMATCH'} ({(iAB[0].cpu().numpy() != i1).sum()} idx)  B->A {'ok' if okBA else 'MISMATCH'} ({(iBA[0].cpu().numpy() != i2).sum()} idx)  fallback rows {tm['fallback_rows']}  tc {tm['tc_ms']:.4f} ms  sel+rerank {tm['rerank_ms']:.4f} ms", flush=True)
    bad += (not okAB) + (not okBA)
    if quick and na >= 500:
        break
if bad:
    print("EXACTNESS FAILURES:", bad)
if not quick:
    a, b, _, _ = synth.make_descriptors(5, 5000, 5000)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    frames = [torch.from_numpy(synth.make_descriptors(100 + f, 2000, 8)[0]).to(dev) for f in range(10)]
    m.pool_reserve(12)
    for f in range(10):
        m.pool_store(f, frames[f])
    m.pool_store(10, ta); m.pool_store(11, tb)
    idx = [(j, i) for i in range(10) for j in range(i + 1, 10)]
    for name, sl, sz, flop in (("cfg5 5000x5000", [(10, 11)], [(5000, 5000)], 2.0 * 5000 * 5000 * 256), ("cfg2 45 x 2000^2", idx, [(2000, 2000)] * 45, 45 * 2.0 * 2000 * 2000 * 256)):
        for _ in range(3):
            m.knn_match_slots(sl, sz, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            m.knn_match_slots(sl, sz, device=dev)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 20
        tm = m.timing()
        print(f"{name}: call {wall*1e3:.3f} ms  tc {tm['tc_ms']:.4f} ms = {flop/(tm['tc_ms']*1e-3)/1e12:.0f} TFLOP/s algorithmic  select+rerank {tm['rerank_ms']:.4f} ms  fallback {tm['fallback_ms']:.4f} ms ({tm['fallback_rows']} rows)  units {tm['units']}", flush=True)
m.close()
