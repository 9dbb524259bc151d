// halo.cu — halo-row exchange between the GPUs of one box without a staging copy: every rank WRITES the feature rows a
// peer asked for straight into that peer's halo buffer through NVLink (peer-mapped device memory, CUDA IPC), one
// kernel for all peers.  Replaces pack (gnnb_gather_rows) + NCCL all-to-all-v; completion is published by a tiny
// collective issued by the host side (partition.py) on the same stream.  No reference counterpart.
#include "common.cuh"

namespace gnnb {

constexpr int MAX_PEERS = 16;
struct PushParams {
    float* peer_base[MAX_PEERS];     // halo buffer of every peer (peer-mapped), nullptr for self / unused
    int64_t peer_row0[MAX_PEERS];    // first row inside that buffer that belongs to this rank
    int64_t seg_start[MAX_PEERS + 1];  // send list is grouped by peer: rows [seg_start[p], seg_start[p+1]) go to peer p
    int64_t stride;                  // multiplicative permutation of the row order (coprime with n_send): every rank writes
                                     // to all of its peers at the same time instead of peer after peer (all ranks hammering
                                     // rank 0's inbound links first: measured 395 GB/s at 4 GPUs)
    int world;
};

// one warp per row: the lanes copy the row as float4 (or float) pieces; the destination is a peer's halo buffer
template <int VEC>
__global__ void __launch_bounds__(256) halo_push_kernel(const PushParams pp, const int32_t* __restrict__ send_idx, int64_t n_send,
                                                        const float* __restrict__ x, int64_t D) {
    const int lane = threadIdx.x & 31;
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (v >= n_send) return;
    const int64_t r = (int64_t)(((uint64_t)v * (uint64_t)pp.stride) % (uint64_t)n_send);   // stride < 2^25, v < 2^31
    int p = 0;
    while (p + 1 < pp.world && r >= pp.seg_start[p + 1]) ++p;       // <= 16 peers: a short scan
    float* dst = pp.peer_base[p] + (size_t)(pp.peer_row0[p] + (r - pp.seg_start[p])) * D;
    const float* src = x + (size_t)__ldg(send_idx + r) * D;
    if (VEC == 4) {
        for (int64_t f = (int64_t)lane * 4; f < D; f += 128)
            *reinterpret_cast<float4*>(dst + f) = __ldg(reinterpret_cast<const float4*>(src + f));
    } else {
        for (int64_t f = lane; f < D; f += 32) dst[f] = __ldg(src + f);
    }
}

}  // namespace gnnb

using namespace gnnb;

extern "C" {

int gnnb_dev_alloc(void** p, int64_t bytes) {
    if (!p || bytes < 0) GNNB_FAIL(GNNB_EINVAL, "dev_alloc: bad argument");
    *p = nullptr;
    GNNB_CUDA(cudaMalloc(p, (size_t)(bytes > 0 ? bytes : 1)));
    return GNNB_OK;
}
int gnnb_dev_free(void* p) {
    if (p) GNNB_CUDA(cudaFree(p));
    return GNNB_OK;
}
int gnnb_ipc_get_handle(void* p, unsigned char* handle64) {
    if (!p || !handle64) GNNB_FAIL(GNNB_EINVAL, "ipc_get_handle: NULL argument");
    cudaIpcMemHandle_t h;
    GNNB_CUDA(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return GNNB_OK;
}
int gnnb_ipc_open_handle(const unsigned char* handle64, void** p) {
    if (!p || !handle64) GNNB_FAIL(GNNB_EINVAL, "ipc_open_handle: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GNNB_CUDA(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
    return GNNB_OK;
}
int gnnb_ipc_close_handle(void* p) {
    if (p) GNNB_CUDA(cudaIpcCloseMemHandle(p));
    return GNNB_OK;
}

int gnnb_halo_push(const int32_t* send_idx_dev, const int64_t* seg_start_host, const void* const* peer_base_host,
                   const int64_t* peer_row0_host, int world, const float* x, int64_t D, void* stream) {
    if (world < 1 || world > MAX_PEERS) GNNB_FAIL(GNNB_EINVAL, "halo_push: world must be in [1,%d]", MAX_PEERS);
    if (!seg_start_host || !peer_base_host || !peer_row0_host) GNNB_FAIL(GNNB_EINVAL, "halo_push: NULL argument");
    if (D <= 0) GNNB_FAIL(GNNB_ESIZE, "halo_push: D must be positive");
    PushParams pp;
    for (int p = 0; p < MAX_PEERS; ++p) { pp.peer_base[p] = nullptr; pp.peer_row0[p] = 0; pp.seg_start[p] = 0; }
    pp.seg_start[MAX_PEERS] = 0;
    pp.world = world;
    for (int p = 0; p < world; ++p) {
        pp.peer_base[p] = (float*)peer_base_host[p];
        pp.peer_row0[p] = peer_row0_host[p];
        pp.seg_start[p] = seg_start_host[p];
        if (seg_start_host[p + 1] > seg_start_host[p] && !peer_base_host[p]) GNNB_FAIL(GNNB_EINVAL, "halo_push: peer %d has rows but no buffer", p);
    }
    pp.seg_start[world] = seg_start_host[world];
    const int64_t n_send = seg_start_host[world];
    if (n_send == 0) return GNNB_OK;
    if (!send_idx_dev || !x) GNNB_FAIL(GNNB_EINVAL, "halo_push: NULL argument");
    pp.stride = 1;
    for (int64_t cand : {1000003LL, 999983LL, 15485863LL, 32452843LL}) {
        if (n_send % cand != 0) { pp.stride = cand % n_send; break; }   // prime not dividing n_send => bijection mod n_send
    }
    if (pp.stride == 0) pp.stride = 1;
    cudaStream_t st = (cudaStream_t)stream;
    bool v4 = D % 4 == 0 && !((uintptr_t)x & 15);
    for (int p = 0; p < world; ++p) if ((uintptr_t)pp.peer_base[p] & 15) v4 = false;
    if (v4) halo_push_kernel<4><<<(unsigned)ceil_div(n_send, 8), 256, 0, st>>>(pp, send_idx_dev, n_send, x, D);
    else halo_push_kernel<1><<<(unsigned)ceil_div(n_send, 8), 256, 0, st>>>(pp, send_idx_dev, n_send, x, D);
    GNNB_LAUNCHED();
    return GNNB_OK;
}

}  // extern "C"
