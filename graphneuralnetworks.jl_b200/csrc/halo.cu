// halo.cu — halo-row exchange between the GPUs of one box without a staging copy: every rank WRITES the feature rows a
// peer asked for straight into that peer's halo buffer through NVLink (peer-mapped device memory, CUDA IPC), one
// kernel for all peers.  Replaces pack (gnnb_gather_rows) + NCCL all-to-all-v; completion is published by a tiny
// collective issued by the host side (partition.py) on the same stream.  No reference counterpart.
#include "common.cuh"

namespace gnnb {

constexpr int MAX_PEERS = 16;
constexpr int PUSH_ROWS = 256;       // rows of one peer's segment a CTA walks through, 8 at a time (one per warp)
struct PushParams {
    float* peer_base[MAX_PEERS];     // halo buffer of every peer (peer-mapped), nullptr for self / unused
    int64_t peer_row0[MAX_PEERS];    // first row inside that buffer that belongs to this rank
    int64_t seg_start[MAX_PEERS + 1];  // send list is grouped by peer: rows [seg_start[p], seg_start[p+1]) go to peer p
    int world;
};

// blockIdx.x = peer, blockIdx.y = chunk of PUSH_ROWS consecutive rows of that peer's segment: CTAs are scheduled x-fastest,
// so a rank feeds all of its peers at once (all ranks hammering rank 0's inbound links first measured 395 GB/s at 4 GPUs),
// while every peer's buffer is written in ascending order by a moving front of a few hundred 8-row windows.  Round 1 /
// early round 2 visited the rows in a pseudo-random order instead: fine for 512 B rows, but at 1 KB rows the scattered
// remote writes fell off a translation cliff (halo_probe: 675 / 566 / 128 GB/s at D = 64 / 128 / 256 with 1.6 M rows).
template <int VEC>
__global__ void __launch_bounds__(256) halo_push_kernel(const PushParams pp, const int32_t* __restrict__ send_idx,
                                                        const float* __restrict__ x, int64_t D) {
    const int p = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t seg0 = pp.seg_start[p], len = pp.seg_start[p + 1] - seg0;
    float* const base = pp.peer_base[p] + (size_t)pp.peer_row0[p] * D;
    for (int64_t c = blockIdx.y; c * PUSH_ROWS < len; c += gridDim.y) {
        const int64_t k1 = (c + 1) * PUSH_ROWS < len ? (c + 1) * PUSH_ROWS : len;
        for (int64_t k = c * PUSH_ROWS + warp; k < k1; k += 8) {
            float* dst = base + (size_t)k * D;
            const float* src = x + (size_t)__ldg(send_idx + seg0 + k) * D;
            if (VEC == 4) {
                for (int64_t f = (int64_t)lane * 4; f < D; f += 128)
                    *reinterpret_cast<float4*>(dst + f) = __ldg(reinterpret_cast<const float4*>(src + f));
            } else {
                for (int64_t f = lane; f < D; f += 32) dst[f] = __ldg(src + f);
            }
        }
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
    int64_t max_seg = 0;
    for (int p = 0; p < world; ++p) {
        const int64_t l = seg_start_host[p + 1] - seg_start_host[p];
        if (l > max_seg) max_seg = l;
    }
    cudaStream_t st = (cudaStream_t)stream;
    bool v4 = D % 4 == 0 && !((uintptr_t)x & 15);
    for (int p = 0; p < world; ++p) if ((uintptr_t)pp.peer_base[p] & 15) v4 = false;
    int64_t chunks = ceil_div(max_seg, PUSH_ROWS);
    if (chunks > 65535) chunks = 65535;                          // the kernel strides over the rest
    const dim3 grid((unsigned)world, (unsigned)chunks);
    if (v4) halo_push_kernel<4><<<grid, 256, 0, st>>>(pp, send_idx_dev, x, D);
    else halo_push_kernel<1><<<grid, 256, 0, st>>>(pp, send_idx_dev, x, D);
    GNNB_LAUNCHED();
    return GNNB_OK;
}

}  // extern "C"
