// ransac.cu — placeholder until the RANSAC kernels land.
#include "bt_common.cuh"
namespace bt { void ransac_destroy(bt_ctx*) {} }
extern "C" int bt_ransac_reserve(bt_ctx*, int, int, int) { bt::set_error("ransac not built yet"); return BT_ERR_UNSUPPORTED; }
extern "C" int bt_ransac_pairs(bt_ctx*, int, const float* const*, const float* const*, const int*, int, float, uint64_t, int32_t*, int32_t*, void*) { bt::set_error("ransac not built yet"); return BT_ERR_UNSUPPORTED; }
