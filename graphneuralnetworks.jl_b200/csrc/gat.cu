// gat.cu — fused GAT edge kernels (placeholder until the fused kernels land in this file).
#include "common.cuh"
using namespace gnnb;
extern "C" {
int gnnb_gat_aggregate(gnnb_graph_t g, const float* Wx, const float* el, const float* er, int64_t C, int64_t H,
                       float slope, float* out, float* alpha, float* seg_max, float* seg_sum, void* stream) {
    GNNB_FAIL(GNNB_EUNSUPPORTED, "gnnb_gat_aggregate: not built yet");
}
int gnnb_gat_aggregate_bwd(gnnb_graph_t g, const float* Wx, const float* el, const float* er, const float* seg_max,
                           const float* seg_sum, const float* dout, int64_t C, int64_t H, float slope, float* dWx,
                           float* del, float* der, void* stream) {
    GNNB_FAIL(GNNB_EUNSUPPORTED, "gnnb_gat_aggregate_bwd: not built yet");
}
}
