// knn.cu — placeholder until the tcgen05 matcher lands (next milestone).
#include "bt_common.cuh"
namespace bt { void matcher_destroy(bt_ctx*) {} }
extern "C" int bt_matcher_reserve(bt_ctx*, int, int, int) { bt::set_error("matcher not built yet"); return BT_ERR_UNSUPPORTED; }
extern "C" int bt_knn_match_pairs(bt_ctx*, int, const bt_desc_view*, const bt_desc_view*, int, int32_t*, float*, int32_t*, float*, void*) { bt::set_error("matcher not built yet"); return BT_ERR_UNSUPPORTED; }
