// Test driver for the C++ host shim (bundletrack_b200/hostcpp/bt_optimizer.hpp) ON A GPU: reads one window, one descriptor pair and
// a few RANSAC pairs from a binary file written by tests/test_host_cpp.py, pushes them through the reference-shaped entry points
// (OptimizerGpu::optimizeFrames with a COLUMN-MAJOR 4x4 like Eigen::Matrix4f, KnnMatcherGpu::knnMatchBothDirections with a
// cv::DMatch-shaped struct, RansacGpu::ransacMultiPairGPU) and writes the results back; the Python side asserts bit-equality with
// the ctypes path.  Plain C++: device memory through the library's own bt_dev_alloc / bt_memcpy_h2d helpers.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "bundletrack_b200/hostcpp/bt_optimizer.hpp"

struct Mat4 { float m[16]; float& operator()(int r, int c) { return m[c * 4 + r]; } float operator()(int r, int c) const { return m[c * 4 + r]; } };   // column-major like Eigen
struct Mat3 { float m[9]; float& operator()(int r, int c) { return m[c * 3 + r]; } float operator()(int r, int c) const { return m[c * 3 + r]; } };
struct EntryJ { unsigned imgIdx_i, imgIdx_j; float pos_i[3], pos_j[3]; };
struct uchar4_ { unsigned char x, y, z, w; };
struct float4_ { float x, y, z, w; };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };

static void rd(FILE* f, void* p, size_t n) { if (n && fread(p, 1, n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); } }
static void wr(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) { std::fprintf(stderr, "short write\n"); std::exit(2); } }
template <class T> static T* upload(const std::vector<T>& h) {
	void* d = nullptr;
	if (bt_dev_alloc(&d, sizeof(T) * (h.empty() ? 1 : h.size())) != BT_OK) { std::fprintf(stderr, "bt_dev_alloc: %s\n", bt_last_error()); std::exit(3); }
	if (!h.empty() && bt_memcpy_h2d(d, h.data(), sizeof(T) * h.size(), nullptr) != BT_OK) { std::fprintf(stderr, "bt_memcpy_h2d: %s\n", bt_last_error()); std::exit(3); }
	return (T*)d;
}

int main(int argc, char** argv) {
	if (argc < 3) return 1;
	FILE* in = std::fopen(argv[1], "rb"); FILE* out = std::fopen(argv[2], "wb");
	if (!in || !out) return 1;
	try {
		// ---- one window through OptimizerGpu::optimizeFrames
		int hdr[4]; float Kf[4];
		rd(in, hdr, sizeof hdr); rd(in, Kf, sizeof Kf);
		const int N = hdr[0], H = hdr[1], W = hdr[2], C = hdr[3];
		std::vector<float*> depths(N); std::vector<float4_*> normals(N); std::vector<uchar4_*> colors(N, nullptr);
		for (int f = 0; f < N; f++) { std::vector<float> d((size_t)H * W); rd(in, d.data(), d.size() * 4); depths[f] = upload(d); }
		for (int f = 0; f < N; f++) { std::vector<float4_> n((size_t)H * W); rd(in, n.data(), n.size() * 16); normals[f] = upload(n); }
		std::vector<EntryJ> corr(C); rd(in, corr.data(), (size_t)C * sizeof(EntryJ));
		std::vector<float> flat((size_t)N * 16); rd(in, flat.data(), flat.size() * 4);
		std::vector<Mat4> poses(N);
		for (int f = 0; f < N; f++) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) poses[f](r, c) = flat[16 * f + 4 * r + c];
		Mat3 K{}; K(0, 0) = Kf[0]; K(1, 1) = Kf[1]; K(0, 2) = Kf[2]; K(1, 2) = Kf[3]; K(2, 2) = 1.f;
		BtSolverConfig cfg;      // the shipped config_nocs.yml values
		OptimizerGpu opt(cfg, 0, N > 2 ? N : 2, C > 0 ? C : 1, H, W, 2);
		std::vector<int> n_match_per_pair;
		{   // Bundler::optimizeGPU's gate (Bundler.cpp:343): too few edges to the new frame -> NO_BA, the poses stay as they are
			std::vector<Mat4> untouched = poses;
			if (opt.optimizeGPU(corr, n_match_per_pair, 10, 10, N, H, W, depths, colors, normals, untouched, K)) return 4;
			for (int f = 0; f < N; f++) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) if (untouched[f](r, c) != poses[f](r, c)) return 5;
		}
		if (!opt.optimizeGPU(corr, n_match_per_pair, 11, 10, N, H, W, depths, colors, normals, poses, K)) return 6;      // = optimizeFrames behind the gate
		saveNewframePose(std::string(argv[2]) + ".pose.txt", poses[N - 1]);
		for (int f = 0; f < N; f++) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) flat[16 * f + 4 * r + c] = poses[f](r, c);
		wr(out, flat.data(), flat.size() * 4);
		{   // the batched form: the same window twice in one call must reproduce the single call
			std::vector<Mat4> p0(N), p1(N);
			std::vector<float> init((size_t)N * 16);
			std::fseek(in, -(long)(init.size() * 4), SEEK_CUR); rd(in, init.data(), init.size() * 4);
			for (int f = 0; f < N; f++) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) p0[f](r, c) = p1[f](r, c) = init[16 * f + 4 * r + c];
			typedef OptimizerGpu::WindowArgs<EntryJ, float4_, Mat4, std::allocator<Mat4>, Mat3> WA;
			std::vector<WA> ws(2);
			ws[0] = WA{ &corr, N, H, W, &depths, &normals, &p0, &K }; ws[1] = WA{ &corr, N, H, W, &depths, &normals, &p1, &K };
			opt.optimizeWindows(ws);
			for (int f = 0; f < N; f++) for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) flat[16 * f + 4 * r + c] = p1[f](r, c);
			wr(out, flat.data(), flat.size() * 4);
		}
		// ---- the two knnMatch calls
		int nab[2]; rd(in, nab, sizeof nab);
		std::vector<float> dA((size_t)nab[0] * 256), dB((size_t)nab[1] * 256);
		rd(in, dA.data(), dA.size() * 4); rd(in, dB.data(), dB.size() * 4);
		float* gA = upload(dA); float* gB = upload(dB);
		KnnMatcherGpu knn(opt.ctx(), nab[0] > nab[1] ? nab[0] : nab[1], 256, 5);
		std::vector<std::vector<DMatch>> ab, ba;
		knn.knnMatchBothDirections(gA, nab[0], 1024, gB, nab[1], 1024, ab, ba);
		knn.knnMatchBothDirections(gA, nab[0], 1024, gB, nab[1], 1024, ab, ba);      // a second call reuses the object's buffers
		for (auto* mm : { &ab, &ba })
			for (auto& row : *mm) {
				int idx[5]; float dist[5];
				for (int j = 0; j < 5; j++) { idx[j] = j < (int)row.size() ? row[j].trainIdx : -1; dist[j] = j < (int)row.size() ? row[j].distance : -1.f; }
				wr(out, idx, sizeof idx); wr(out, dist, sizeof dist);
			}
		// ---- ransacMultiPairGPU
		int npairs; float thr; rd(in, &npairs, 4); rd(in, &thr, 4);
		std::vector<float4_*> pa(npairs), pb(npairs); std::vector<int> np(npairs);
		int maxp = 1;
		for (int p = 0; p < npairs; p++) {
			rd(in, &np[p], 4);
			std::vector<float4_> a(np[p]), b(np[p]);
			rd(in, a.data(), a.size() * 16); rd(in, b.data(), b.size() * 16);
			pa[p] = upload(a); pb[p] = upload(b);
			if (np[p] > maxp) maxp = np[p];
		}
		RansacGpu rs(opt.ctx(), npairs > 0 ? npairs : 1, maxp, 2000);
		std::vector<std::vector<int>> inl;
		rs.ransacMultiPairGPU(pa, pb, np, 2000, thr, inl);
		rs.ransacMultiPairGPU(pa, pb, np, 2000, thr, inl);
		for (int p = 0; p < npairs; p++) { const int c = (int)inl[p].size(); wr(out, &c, 4); wr(out, inl[p].data(), (size_t)c * 4); }
	} catch (const std::exception& e) {
		std::fprintf(stderr, "shim_gpu: %s\n", e.what());
		return 4;
	}
	std::fclose(in); std::fclose(out);
	return 0;
}
