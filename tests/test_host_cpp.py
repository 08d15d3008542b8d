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
