"""The C++ host shim (reference-shaped OptimizerGpu::optimizeFrames / ransacMultiPairGPU / knnMatch / depth front end on top of the C-ABI) must compile
against a minimal Eigen-like matrix type and link with the in-tree library."""
import os
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_compiles_and_links(tmp_path):
    src = tmp_path / "shim_test.cpp"
    src.write_text(textwrap.dedent(r'''
        #include "bundletrack_b200/hostcpp/bt_optimizer.hpp"
        #include <cstdio>
        struct Mat4 { float m[16]; float& operator()(int r, int c) { return m[c * 4 + r]; } float operator()(int r, int c) const { return m[c * 4 + r]; } };  // column-major like Eigen
        struct Mat3 { float m[9]; float operator()(int r, int c) const { return m[c * 3 + r]; } };
        struct EntryJ { unsigned imgIdx_i, imgIdx_j; float pos_i[3], pos_j[3]; };
        struct uchar4_ { unsigned char x, y, z, w; }; struct float4_ { float x, y, z, w; };
        int main() {
            try {
                BtSolverConfig cfg;
                OptimizerGpu opt(cfg);            // throws without a GPU: the product has no CPU fallback
                std::vector<EntryJ> corr; std::vector<int> nm; std::vector<float*> d; std::vector<uchar4_*> c; std::vector<float4_*> n; std::vector<Mat4> poses; Mat3 K{};
                opt.optimizeFrames(corr, nm, 0, 480, 640, d, c, n, poses, K);
                // the other two call sites of the path and the frame front end instantiate too
                struct DMatch { int queryIdx, trainIdx; float distance; };
                std::vector<std::vector<DMatch>> ab, ba;
                knnMatchBothDirections(opt.ctx(), (const float*)nullptr, 0, 1024, (const float*)nullptr, 0, 1024, 256, 5, ab, ba);
                std::vector<float4_*> pa, pb; std::vector<int> np; std::vector<std::vector<int>> inl;
                ransacMultiPairGPU(opt.ctx(), pa, pb, np, 2000, 0.005f, inl);
                processDepthAndNormals(opt.ctx(), (const float*)nullptr, (float*)nullptr, (float4_*)nullptr, (float4_*)nullptr, 480, 640, K, BtDepthConfig());
            } catch (const std::exception& e) { std::printf("caught: %s\n", e.what()); return 0; }
            return 0;
        }
    '''))
    exe = tmp_path / "shim_test"
    lib_dir = os.path.join(ROOT, "bundletrack_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-I", ROOT, str(src), "-o", str(exe), "-L", lib_dir, "-lbundletrack_b200", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr


import numpy as np
import pytest


@pytest.mark.gpu
def test_cpp_shim_runs_on_gpu_and_equals_ctypes_path(tmp_path, cuda_device):
    """The reference-shaped C++ entry points with DATA on a GPU: OptimizerGpu::optimizeFrames on a 10-keyframe x 2000-correspondence
    window through a column-major Mat4 (the Bundler.cpp:350-351 drop-in), the batched optimizeWindows, KnnMatcherGpu and RansacGpu -
    each bit-identical to the ctypes path the other tests use."""
    import torch
    from bundletrack_b200 import synth
    from bundletrack_b200.optimizer import OptimizerGpu, SolveWindow
    from bundletrack_b200.matcher import KnnMatcher, Ransac
    w = synth.make_window(33, n_frames=10, n_corr=2000)
    a, b, _, _ = synth.make_descriptors(12, 700, 900)
    rcases = [synth.make_ransac_case(400 + k, n) for k, n in enumerate((40, 500, 1500))]
    thr = np.float32(0.005)
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        np.array([w.n_frames, w.H, w.W, len(w.corr)], np.int32).tofile(f)
        np.asarray(w.K, np.float32).tofile(f)
        np.ascontiguousarray(w.depth, np.float32).tofile(f)
        np.ascontiguousarray(w.normal, np.float32).tofile(f)
        w.corr.tofile(f)
        np.ascontiguousarray(w.poses_init, np.float32).tofile(f)
        np.array([len(a), len(b)], np.int32).tofile(f)
        a.tofile(f); b.tofile(f)
        np.array([len(rcases)], np.int32).tofile(f); np.array([thr], np.float32).tofile(f)
        for A4, B4, _ in rcases:
            np.array([len(A4)], np.int32).tofile(f); A4.tofile(f); B4.tofile(f)
    exe = tmp_path / "shim_gpu"
    lib_dir = os.path.join(ROOT, "bundletrack_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", ROOT, os.path.join(ROOT, "tests", "cpp", "shim_gpu_main.cpp"), "-o", str(exe),
                           "-L", lib_dir, "-lbundletrack_b200", f"-Wl,-rpath,{lib_dir}"])
    outp = tmp_path / "out.bin"
    run = subprocess.run([str(exe), str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    raw = np.fromfile(outp, np.uint8)
    N = w.n_frames
    o = 0
    def take(dtype, count):
        nonlocal o
        v = raw[o:o + count * np.dtype(dtype).itemsize].view(dtype); o += count * np.dtype(dtype).itemsize
        return v
    poses_cpp = take(np.float32, N * 16).reshape(N, 4, 4)
    poses_cpp_batch = take(np.float32, N * 16).reshape(N, 4, 4)
    # ctypes path
    dev = cuda_device
    depth = [torch.from_numpy(np.ascontiguousarray(w.depth[k])).to(dev) for k in range(N)]
    normal = [torch.from_numpy(np.ascontiguousarray(w.normal[k])).to(dev) for k in range(N)]
    opt = OptimizerGpu(None, max_windows=2, max_frames=10, max_corr=2000)
    win = SolveWindow(w.corr, w.H, w.W, depth, normal, w.poses_init, w.K)
    ref = opt.optimizeWindows([win])[0]
    ref_batch = opt.optimizeWindows([win, win])[1]       # (a batch picks another tile size: same poses to ~1e-6, not bit for bit)
    opt.close()
    assert np.array_equal(poses_cpp, ref)
    assert np.array_equal(poses_cpp_batch, ref_batch)
    # saveNewframePose (Bundler::saveNewframeResult's pose record): ob_in_cam of the newest frame, as text
    from bundletrack_b200 import policy
    assert open(str(outp) + ".pose.txt").read() == policy.pose_text(ref[N - 1])
    assert synth.pose_errors(ref, w.poses_gt)[0] < synth.pose_errors(w.poses_init, w.poses_gt)[0]
    m = KnnMatcher(max_pairs=1, max_feats=1024)
    iAB, dAB, iBA, dBA = m.knn_match_pairs([(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))])
    m.close()
    for n, ii, dd in ((len(a), iAB[0], dAB[0]), (len(b), iBA[0], dBA[0])):
        rec = take(np.uint8, n * 40).reshape(n, 40)
        assert np.array_equal(rec[:, :20].copy().view(np.int32).reshape(n, 5), ii.cpu().numpy())
        assert np.array_equal(rec[:, 20:].copy().view(np.float32).reshape(n, 5), dd.cpu().numpy())
    r = Ransac(max_pairs=4, max_pts=2048, max_trials=2000)
    ids = r.ransac_pairs([torch.from_numpy(c[0]).to(dev) for c in rcases], [torch.from_numpy(c[1]).to(dev) for c in rcases], 2000, float(thr))
    r.close()
    for p in range(len(rcases)):
        cnt = int(take(np.int32, 1)[0])
        assert np.array_equal(take(np.int32, cnt), ids[p].cpu().numpy())
    assert o == len(raw)
