#!/usr/bin/env python
"""bench.py — pose-graph windows/sec (10 keyframes x 2000 correspondences, 640x480 synthetic RGB-D) on N B200s.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` (torchrun for N>1) prints ONE JSON
line on rank 0.  One "step" = one pass of the hot path over one batch of `--windows` independent tracking windows per GPU
(BASELINE.json configs[1] replicated; configs[3] is 32 windows/GPU x 8 GPUs), i.e. weak scaling.

  value     windows/s, whole job, inputs (frame maps, correspondences, poses, tables) already resident in HBM:
            the three kernels of bt_solve_run timed with CUDA events on the launching stream.
  e2e       the same metric through the reference-facing call (OptimizerGpu.optimizeWindows -> bt_solve_windows):
            per step the host correspondences + poses + window tables go host->device and the poses come back, exactly
            the arguments OptimizerGpu::optimizeFrames takes from the host; depth/normal maps are device-resident
            Frame members in the reference API (Frame::_depth_gpu/_normal_gpu) and are passed as device pointers.
  roofline  dominant kernel = k_solve; achieved = algorithmic bytes (iters*(N*npix*32 + C*32) + 2*N*64
            per window) / CUDA-event duration of that kernel; peak = MEASURED_PEAKS.json hbm_gbs (fallback 6650).
  cpu_baseline  oracle/ (CPU restatement, "port": the reference has no CPU optimizer) on the host cores, bounded sample.
  --impl reference   the reference's OWN CUDA kernels + host-glue allocation pattern (oracle/_ref, built verbatim from
            /root/reference) called once per window like Bundler::optimizeGPU does; plus the CPU port beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=32, help="windows per GPU per step (BASELINE configs[3]: 256 over 8 GPUs)")
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--corr", type=int, default=2000)
    ap.add_argument("--scenes", type=int, default=4, help="distinct rendered scenes per GPU; windows cycle through them with their own pose noise and their own copy of the frame maps")
    ap.add_argument("--cpu-windows", type=int, default=0, help="cpu_baseline sample size (0 = 2 per core)")
    ap.add_argument("--ref-windows", type=int, default=8, help="windows per step for --impl reference")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def make_batch(args, rank, dev):
    """`windows` windows per GPU: `scenes` rendered scenes, every window gets its OWN device copy of the frame maps (so
    the batch footprint, windows*frames*6.1 MB, is far larger than the 126 MB L2) and its own initial-pose noise."""
    import torch
    from bundletrack_b200 import synth
    from bundletrack_b200.optimizer import SolveWindow
    scenes = [synth.make_window(1000 * rank + s, n_frames=args.frames, n_corr=args.corr) for s in range(min(args.scenes, args.windows))]
    wins, host = [], []
    # every window's correspondences live in ONE page-locked host block (the contract's "inputs from pinned host memory"): the
    # library then lets the copy engine read them in place instead of staging them through its own pinned buffer
    n_ent = sum(len(scenes[k % len(scenes)].corr) for k in range(args.windows))
    pinned = torch.empty(max(n_ent, 1) * 32, dtype=torch.uint8).pin_memory()
    corr_all = pinned.numpy().view(synth.ENTRYJ_DTYPE)
    make_batch.keep = pinned
    c0 = 0
    for k in range(args.windows):
        sc = scenes[k % len(scenes)]
        corr_k = corr_all[c0:c0 + len(sc.corr)]
        corr_k[:] = sc.corr
        c0 += len(sc.corr)
        rng = np.random.default_rng(77 + 1000 * rank + k)
        poses = sc.poses_gt.copy()
        for f in range(1, sc.n_frames):
            poses[f] = sc.poses_gt[f] @ synth.se3(synth.so3_exp(rng.normal(0, np.deg2rad(1.0), 3)), rng.normal(0, 0.003, 3))
        depth = [torch.from_numpy(sc.depth[f]).to(dev) for f in range(sc.n_frames)]
        normal = [torch.from_numpy(sc.normal[f]).to(dev) for f in range(sc.n_frames)]
        wins.append(SolveWindow(corr_k, sc.H, sc.W, depth, normal, poses.astype(np.float32), sc.K))
        host.append((sc, poses.astype(np.float32)))
    return wins, host


def matcher_microbench(dev, stream):
    """Secondary evidence for the matcher half of the path (BASELINE configs[4]/[1]): kernel times from CUDA events inside
    the library; tensor fraction = executed bf16 flops (both directions) / measured cuBLAS bf16 peak."""
    import torch
    from bundletrack_b200 import synth
    from bundletrack_b200.matcher import KnnMatcher
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tf_peak, src = float(json.load(f)["bf16_tflops"]), "measured"
    except Exception:
        tf_peak, src = 1590.0, "fallback"
    m = KnnMatcher(max_pairs=48, max_feats=5120, stream=stream)
    m.enable_timing(True)
    out = {"peak_bf16_tflops": tf_peak, "peak_source": src}
    a, b, _, _ = synth.make_descriptors(5, 5000, 5000)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    frames = [torch.from_numpy(synth.make_descriptors(100 + f, 2000, 8)[0]).to(dev) for f in range(10)]
    pairs = [(frames[j], frames[i]) for i in range(10) for j in range(i + 1, 10)]
    for name, prs, flop in (("cfg5_5000x5000", [(ta, tb)], 2 * 2.0 * 5000 * 5000 * 256), ("cfg2_45pairs_x2000", pairs, 45 * 2 * 2.0 * 2000 * 2000 * 256)):
        for _ in range(3):
            m.knn_match_pairs(prs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            m.knn_match_pairs(prs)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 10
        tm = m.timing()
        out[name] = {"pairs": len(prs), "call_ms": wall * 1e3, "pairs_per_s": len(prs) / wall, "tc_kernel_ms": tm["tc_ms"], "rerank_ms": tm["rerank_ms"],
                     "fallback_ms": tm["fallback_ms"], "fallback_rows": tm["fallback_rows"], "tc_tflops_executed": flop / (tm["tc_ms"] * 1e-3) / 1e12,
                     "tc_frac_of_bf16_peak": flop / (tm["tc_ms"] * 1e-3) / 1e12 / tf_peak}
    m.close()
    try:      # the whole device-resident chain of the matcher side: kNN -> prune -> mutual -> RANSAC (2000 trials) -> EntryJ, 45 pairs of a 10-frame window
        from bundletrack_b200.matcher import MatchPipeline
        w = synth.make_window(77, n_frames=10, n_corr=10)
        fr = synth.make_feature_frames(w, 2000, seed=77, n_surface=12000)
        devf = [{"kpts": torch.from_numpy(fr[k]["kpts"]).to(dev), "desc": torch.from_numpy(fr[k]["desc"]).to(dev), "depth": torch.from_numpy(w.depth[k]).to(dev),
                 "normal": torch.from_numpy(w.normal[k]).to(dev), "pose": w.poses_init[k], "id": k, "window_index": k} for k in range(10)]
        mp = MatchPipeline(None, max_pairs=48, max_feats=2048, stream=stream)
        prs = [(devf[j], devf[i]) for i in range(10) for j in range(i + 1, 10)]
        for _ in range(3):
            ent, n_ent, off = mp.match_pairs(prs, w.H, w.W, w.K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            ent, n_ent, off = mp.match_pairs(prs, w.H, w.W, w.K)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 10
        out["pipeline_45pairs"] = {"call_ms": wall * 1e3, "pairs_per_s": 45 / wall, "feats_per_frame": int(np.mean([len(f["kpts"]) for f in fr])),
                                   "entries_out": int(np.sum(n_ent)), "note": "bt_match_pairs incl. the D2H of the EntryJ list"}
        mp.close()
    except Exception as e:
        out["pipeline_error"] = repr(e)
    return out


def next_row_microbench(opt, wins, dev, stream, args):
    """Secondary numbers for SURVEY.md 8f rank 1 (NOT the headline): (1) the same batch with every keyframe's quarter-res maps
    taken from the frame cache (built once per keyframe by bt_frame_cache_store) instead of rebuilt inside the call;
    (2) the fused depth front end (erode + 2x filter + points + normals) on 640x480 frames against its 40 B/pixel roofline."""
    import torch
    from bundletrack_b200 import synth
    from bundletrack_b200.optimizer import SolveWindow
    from bundletrack_b200.frontend import FrameFrontEnd
    out = {}
    try:
        peak, _ = hbm_peak()
        N = wins[0].n_frames
        nF = len(wins) * N
        opt.reserve_frame_cache(nF, wins[0].H, wins[0].W)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        slots = list(range(nF))
        dps = [d for w in wins for d in w.depths]
        nps = [n for w in wins for n in w.normals]
        for _ in range(2):
            opt.store_frames(slots, dps, nps, wins[0].H, wins[0].W, wins[0].K)
        ev[0].record()
        for _ in range(5):
            opt.store_frames(slots, dps, nps, wins[0].H, wins[0].W, wins[0].K)
        ev[1].record()
        cw = [SolveWindow(w.corr, w.H, w.W, None, None, w.poses, w.K, cache_slots=slots[i * N:(i + 1) * N]) for i, w in enumerate(wins)]
        opt.stage(cw)
        for _ in range(args.warmup):
            opt.run()
        ev[2].record()
        for _ in range(args.steps):
            opt.run()
        ev[3].record()
        torch.cuda.synchronize()
        ms = ev[2].elapsed_time(ev[3]) / args.steps
        out["frame_cache"] = {"value": len(wins) / (ms * 1e-3), "unit": "windows/s", "ms_per_step": ms, "store_us_per_frame": ev[0].elapsed_time(ev[1]) / 5 / nF * 1e3,
                              "note": "keyframe maps built once by bt_frame_cache_store (outside the step); poses identical to the headline path"}
        # one window at a time: the reference's actual calling pattern (one optimizeFrames per tracked frame)
        one = [wins[0]]
        for _ in range(5):
            opt.optimizeWindows(one)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            opt.optimizeWindows(one)
        torch.cuda.synchronize()
        tm1 = opt.timing_ms()
        out["single_window"] = {"e2e_ms": (time.perf_counter() - t0) / 50 * 1e3, "kernel_ms": {"prep": tm1["prep"], "solve": tm1["solve"]},
                                "note": "bt_solve_windows on ONE 10-keyframe x 2000-correspondence window, host buffers in/out"}
        fe = FrameFrontEnd(ctx=opt.ctx, stream=stream)
        nfr, H, W = 64, 480, 640
        raws = [torch.from_numpy(synth.make_raw_depth(s, H, W)[0]).to(dev) for s in range(4)]
        tin = [raws[i % 4].clone() for i in range(nfr)]
        dout = [torch.empty((H, W), device=dev) for _ in range(nfr)]
        nout = [torch.empty((H, W, 4), device=dev) for _ in range(nfr)]
        xout = [torch.empty((H, W, 4), device=dev) for _ in range(nfr)]
        for _ in range(3):
            fe.process(tin, H, W, synth.NOCS_K, dout, nout, xout)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fe.process(tin, H, W, synth.NOCS_K, dout, nout, xout)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10 * 1e-3
        gbs = nfr * H * W * 40 / t / 1e9
        out["frontend"] = {"frames_per_s": nfr / t, "us_per_frame": t / nfr * 1e6, "achieved_GBps": gbs, "peak_GBps": peak, "frac": gbs / peak,
                           "algorithmic_bytes_per_pixel": 40, "sample": f"{nfr} frames 640x480 per call, 4+16+16 B written and 4 B read per pixel"}
    except Exception as e:      # secondary evidence must never take the headline down
        out["next_row_error"] = repr(e)
    return out


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same bench
    command (profiles/traffic_r*.json, written by scripts/ncu_summary.py); None when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def cpu_baseline(host, n_windows, check_against=None):
    """Oracle A (float build) on the host cores: threads over windows (the C call releases the GIL).  `check_against`: the GPU
    poses of window 0 - the oracle's own result for that window doubles as the parity check of this very run."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    n = n_windows or 2 * cores
    oracle.build()
    jobs = [host[k % len(host)] for k in range(n)]
    def run(j):
        sc, poses = j
        return oracle.solve_window(sc.depth, sc.normal, sc.K, sc.corr, poses)
    run(jobs[0])
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        res = list(ex.map(run, jobs))
    dt = time.perf_counter() - t0
    out = {"value": n / dt, "unit": "windows/s", "cores": cores, "kind": "port",
           "sample": f"{n} windows of the bench workload, oracle/solver_oracle.c (fp32), {cores} threads, {dt:.1f} s"}
    if check_against is not None:
        from bundletrack_b200 import synth
        r, t = synth.pose_errors(check_against, res[0])
        out["gpu_vs_port_window0"] = {"rot_rad": r, "trans_m": t}
    return out


def main():
    args = parse()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.impl == "reference" and rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    windows_cfg = args.windows
    if args.impl == "reference":      # the reference arm only touches its sample: do not keep 1.9 GB of unused maps resident next to its cudaMalloc/cudaFree pattern
        args.windows = min(args.windows, args.ref_windows)
    wins, host = make_batch(args, rank, dev)
    args.windows = windows_cfg
    N, C = args.frames, args.corr
    npix = (640 // 4) * (480 // 4)
    workload = f"{args.windows} windows/GPU x ({N} keyframes, {C} corr, 640x480 -> 160x120 cache, 7 GN x 5 PCG), BASELINE configs[1] batched as configs[3]"
    base = {"metric": "pose-graph windows/sec (10 KF x 2k corr)", "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "windows_per_gpu": args.windows, "frames": N, "corr": C, "gn_iters": 7, "pcg_iters": 5,
                       "l2": f"inputs larger than L2: {args.windows * N * 6.1:.0f} MB of frame maps per GPU, each window its own copy", "parallelism": f"windows sharded, {world} rank(s), no data-path collective"}}

    if args.impl == "reference":
        import oracle
        try:
            oracle.ref_lib()
        except Exception as e:      # oracle/_ref did not travel: the CPU port of the same algorithm stands in (kind "port")
            cb = cpu_baseline(host, args.cpu_windows)
            out = dict(base, impl="reference", value=cb["value"], ms_per_step=None, gpu_launches=0,
                       e2e={"value": cb["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, cpu_baseline=cb,
                       note=f"oracle/_ref unavailable ({e}); timed the CPU port instead")
            print(json.dumps(out))
            if world > 1:
                dist.barrier(); dist.destroy_process_group()
            return
        n = min(args.ref_windows, len(wins))
        def step():
            for k in range(n):
                w = wins[k]
                oracle.ref_optimize_frames(w.depths, w.normals, w.H, w.W, w.K, w.corr, w.poses)
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler(local); sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        val = n * args.steps / dt
        cb = cpu_baseline(host, args.cpu_windows)
        out = dict(base, impl="reference", value=val, ms_per_step=dt / args.steps * 1e3, gpu_launches=0,
                   e2e={"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   cpu_baseline={"value": val, "unit": "windows/s", "cores": 1, "kind": "reference",
                                 "sample": f"{n} windows per step, one optimizeFrames-equivalent call each through the reference's OWN kernels + allocation pattern "
                                           "(oracle/_ref, 1 host thread driving GPU 0): the reference has no CPU optimizer (SURVEY.md D5), its implementation of this path IS CUDA"},
                   cpu_port=cb, clocks=clocks)
        out["config"] = dict(base["config"], reference_sample=f"{n} windows per step through the reference's kernels + allocation pattern")
        print(json.dumps(out))
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    from bundletrack_b200.optimizer import OptimizerGpu
    stream = torch.cuda.current_stream().cuda_stream
    opt = OptimizerGpu(None, device=local, max_windows=args.windows, max_frames=max(N, 2), max_corr=max(C, 1), stream=stream)
    opt.enable_timing(True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: resident inputs, kernels only
    opt.stage(wins)
    for _ in range(args.warmup):
        opt.run()
    sync_all()
    sampler = ClockSampler(local); sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ms = []
    ev0.record()
    for _ in range(args.steps):
        opt.run()
        k_ms.append(None)
    ev1.record()
    sync_all()
    elapsed = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    tm = opt.timing_ms()                      # per-kernel device time of the last step
    stats = opt.stats()
    poses = opt.fetch()
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        from bundletrack_b200.sharding import gather_poses   # NCCL only gathers the results (windows are sharded per rank)
        # every rank holds `windows` windows; global ids are rank-strided like shard_indices
        gather_poses(poses, args.windows * world, [N] * (args.windows * world), rank, world, device=dev)
    ms_step = float(elapsed.item()) / args.steps
    value = world * args.windows / (ms_step * 1e-3)

    # ---- e2e: host buffers in, host poses out, every step
    for _ in range(args.warmup):
        opt.optimizeWindows(wins)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_poses = opt.optimizeWindows(wins)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    clocks = sampler.stop()      # sampled across both timed regions (value + e2e)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_val = world * args.windows * args.steps / float(dt.item())
    h2d = sum(len(w.corr) * 32 + w.n_frames * 64 + w.n_frames * 20 + 120 + 45 * 8 + 46 * 12 for w in wins)
    d2h = sum(w.n_frames * 64 for w in wins)

    matcher = None
    extras = {}
    if rank == 0:
        matcher = matcher_microbench(dev, stream)
        extras = next_row_microbench(opt, wins, dev, stream, args)
    if rank == 0:
        peak, peak_src = hbm_peak()
        alg_bytes = args.windows * (7 * (N * npix * 32 + C * 32) + 2 * N * 64)
        ach = alg_bytes / (tm["solve"] * 1e-3) / 1e9
        cb = cpu_baseline(host, args.cpu_windows, check_against=out_poses[0])
        out = dict(base, value=value, ms_per_step=ms_step, gpu_launches=int(stats["n_kernel_launches"]) * args.steps,
                   e2e={"value": e2e_val, "unit": "windows/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                   roofline={"bound": "hbm", "kernel": "k_solve", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic("k_solve"),
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": tm},
                   cpu_baseline=cb, clocks=clocks,
                   matcher=matcher, **extras,
                   solver_stats=stats)
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    opt.close()


if __name__ == "__main__":
    main()
