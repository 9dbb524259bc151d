// plan.cu — the graph plan: COO (Julia 1-based Int64/Int32, host or device) -> int32 0-based device
// COO + CSR-by-target (+ lazily CSR-by-source), stable in COO order.
//
// Reference counterparts: the COO GNNGraph `(s,t)` (GNNGraphs/src/gnngraph.jl:108-117), index-range
// asserts of to_coo (GNNGraphs/src/convert.jl:49-54), add_self_loops (GNNGraphs/src/transform.jl:12-28).
// The reference has no CSR type and no GPU edge sort (GNNGraphs/ext/GNNGraphsCUDAExt.jl:24-30 sorts on
// the CPU); it rebuilds a CSC from COO on every fused CPU call (GNNGraphs/src/query.jl:227).
#include "common.cuh"
#include <cub/cub.cuh>

namespace gnnb {

static int g_chunk_default = 128;

// ---- kernels ------------------------------------------------------------------------------------
template <typename T>
__global__ void convert_index_kernel(const T* __restrict__ in, int64_t n, int64_t base, int64_t limit,
                                     int32_t* __restrict__ out, int* __restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t v = (int64_t)in[i] - base;
    if (v < 0 || v >= limit) {
        *bad = 1;
        v = 0;
    }
    out[i] = (int32_t)v;
}

__global__ void iota_kernel(int32_t* out, int64_t n, int32_t start) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = start + (int32_t)i;
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx,
                                  int64_t n, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

// rowptr from the sorted row array: rowptr[r] = first position whose row >= r
__global__ void rowptr_kernel(const int32_t* __restrict__ row, int64_t E, int32_t nrows,
                              int32_t* __restrict__ rowptr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > E) return;
    int32_t lo = (i == 0) ? -1 : row[i - 1];
    int32_t hi = (i == E) ? nrows : row[i];
    for (int32_t r = lo + 1; r <= hi; ++r) rowptr[r] = (int32_t)i;
}

__global__ void long_rows_kernel(const int32_t* __restrict__ rowptr, int32_t nrows, int32_t chunk,
                                 int32_t* __restrict__ list, int32_t* __restrict__ count) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    if (rowptr[r + 1] - rowptr[r] > chunk) list[atomicAdd(count, 1)] = (int32_t)r;
}

__global__ void invdeg_kernel(const int32_t* __restrict__ rowptr, int32_t nrows, float* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int d = rowptr[r + 1] - rowptr[r];
    out[r] = 1.0f / (float)(d > 0 ? d : 1);
}

// CSR of add_self_loops(g) from the CSR of g: row r gains one trailing entry (the appended loop
// sorts last inside its row because the sort is stable and loops come after the originals).
__global__ void selfloop_edges_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                                      const int32_t* __restrict__ eid, int64_t E,
                                      int32_t* __restrict__ nrow, int32_t* __restrict__ ncol,
                                      int32_t* __restrict__ neid) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int32_t r = row[e];
    int64_t pos = e + r;
    nrow[pos] = r;
    ncol[pos] = col[e];
    neid[pos] = eid[e];
}
__global__ void selfloop_rows_kernel(const int32_t* __restrict__ rowptr, int32_t n, int32_t E,
                                     int32_t* __restrict__ nrowptr, int32_t* __restrict__ nrow,
                                     int32_t* __restrict__ ncol, int32_t* __restrict__ neid) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    nrowptr[r] = rowptr[r] + (int32_t)r;
    if (r < n) {
        int64_t pos = (int64_t)rowptr[r + 1] + r;  // last slot of the new row r
        nrow[pos] = (int32_t)r;
        ncol[pos] = (int32_t)r;
        neid[pos] = E + (int32_t)r;
    }
}

// ---- helpers ------------------------------------------------------------------------------------
static void free_csr(Csr& c) {
    cudaFree(c.rowptr); cudaFree(c.col); cudaFree(c.row); cudaFree(c.eid);
    cudaFree(c.long_rows); cudaFree(c.invdeg); cudaFree(c.items); cudaFree(c.es);
    c = Csr();
}

static int alloc_csr(Csr& c, int64_t E, int32_t nrows, int32_t ncols, int32_t chunk) {
    c.nrows = nrows;
    c.ncols = ncols;
    GNNB_CUDA(cudaMalloc(&c.rowptr, sizeof(int32_t) * ((size_t)nrows + 1)));
    GNNB_CUDA(cudaMalloc(&c.col, sizeof(int32_t) * (size_t)(E > 0 ? E : 1)));
    GNNB_CUDA(cudaMalloc(&c.row, sizeof(int32_t) * (size_t)(E > 0 ? E : 1)));
    GNNB_CUDA(cudaMalloc(&c.eid, sizeof(int32_t) * (size_t)(E > 0 ? E : 1)));
    GNNB_CUDA(cudaMalloc(&c.long_rows, sizeof(int32_t) * (size_t)(E / chunk + 2)));
    return GNNB_OK;
}

static int find_long_rows(Csr& c, int32_t chunk, cudaStream_t st) {
    int32_t* d_count = nullptr;
    GNNB_CUDA(cudaMalloc(&d_count, sizeof(int32_t)));
    GNNB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int32_t), st));
    if (c.nrows > 0) {
        long_rows_kernel<<<(unsigned)ceil_div(c.nrows, 256), 256, 0, st>>>(c.rowptr, c.nrows, chunk,
                                                                           c.long_rows, d_count);
        GNNB_LAUNCHED();
    }
    GNNB_CUDA(cudaMemcpyAsync(&c.n_long, d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GNNB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_count);
    return GNNB_OK;
}

int ensure_csr(gnnb_graph* g, bool transposed, cudaStream_t st) {
    Csr& c = transposed ? g->by_src : g->by_dst;
    if (c.built) return GNNB_OK;
    std::lock_guard<std::mutex> lock(g->mu);
    if (c.built) return GNNB_OK;
    const int64_t E = g->E;
    const int32_t nrows = transposed ? g->n_src : g->n_dst;
    const int32_t ncols = transposed ? g->n_dst : g->n_src;
    const int32_t* keys = transposed ? g->coo_src : g->coo_dst;
    const int32_t* other = transposed ? g->coo_dst : g->coo_src;
    GNNB_TRY(alloc_csr(c, E, nrows, ncols, g->chunk));
    if (E > 0) {
        int32_t* iota = nullptr;
        GNNB_CUDA(cudaMalloc(&iota, sizeof(int32_t) * (size_t)E));
        iota_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(iota, E, 0);
        GNNB_LAUNCHED();
        int end_bit = 1;
        while (end_bit < 31 && ((int64_t)1 << end_bit) < (int64_t)nrows) ++end_bit;
        size_t tmp_bytes = 0;
        GNNB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, c.row, iota, c.eid, (int)E, 0,
                                                  end_bit, st));
        void* tmp = nullptr;
        GNNB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
        GNNB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, c.row, iota, c.eid, (int)E, 0,
                                                  end_bit, st));
        g_launches.fetch_add(4, std::memory_order_relaxed);  // histogram + onesweep passes (library kernels)
        gather_i32_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(other, c.eid, E, c.col);
        GNNB_LAUNCHED();
        GNNB_CUDA(cudaStreamSynchronize(st));
        cudaFree(tmp);
        cudaFree(iota);
    }
    rowptr_kernel<<<(unsigned)ceil_div(E + 1, 256), 256, 0, st>>>(c.row, E, nrows, c.rowptr);
    GNNB_LAUNCHED();
    GNNB_TRY(find_long_rows(c, g->chunk, st));
    c.built = true;
    return GNNB_OK;
}

int ensure_invdeg(gnnb_graph* g, Csr& c, cudaStream_t st) {
    if (c.invdeg) return GNNB_OK;
    std::lock_guard<std::mutex> lock(g->mu);
    if (c.invdeg) return GNNB_OK;
    float* p = nullptr;
    GNNB_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)(c.nrows > 0 ? c.nrows : 1)));
    if (c.nrows > 0) {
        invdeg_kernel<<<(unsigned)ceil_div(c.nrows, 256), 256, 0, st>>>(c.rowptr, c.nrows, p);
        GNNB_LAUNCHED();
    }
    c.invdeg = p;
    return GNNB_OK;
}

int ensure_ws(gnnb_graph* g, size_t bytes) {
    if (g->ws_bytes >= bytes) return GNNB_OK;
    if (g->ws) { cudaDeviceSynchronize(); cudaFree(g->ws); g->ws = nullptr; g->ws_bytes = 0; }
    GNNB_CUDA(cudaMalloc(&g->ws, bytes));
    g->ws_bytes = bytes;
    return GNNB_OK;
}
int ensure_ws2(gnnb_graph* g, size_t bytes) {
    if (g->ws2_bytes >= bytes) return GNNB_OK;
    if (g->ws2) { cudaDeviceSynchronize(); cudaFree(g->ws2); g->ws2 = nullptr; g->ws2_bytes = 0; }
    GNNB_CUDA(cudaMalloc(&g->ws2, bytes));
    g->ws2_bytes = bytes;
    return GNNB_OK;
}

static int convert_indices(const void* p, int64_t n, int index_bytes, int index_base, int64_t limit,
                           int on_device, int32_t* out, int* d_bad, cudaStream_t st) {
    if (n == 0) return GNNB_OK;
    const void* dev = p;
    void* staged = nullptr;
    if (!on_device) {
        GNNB_CUDA(cudaMalloc(&staged, (size_t)n * index_bytes));
        GNNB_CUDA(cudaMemcpyAsync(staged, p, (size_t)n * index_bytes, cudaMemcpyHostToDevice, st));
        dev = staged;
    }
    unsigned blocks = (unsigned)ceil_div(n, 256);
    if (index_bytes == 8)
        convert_index_kernel<int64_t><<<blocks, 256, 0, st>>>((const int64_t*)dev, n, index_base, limit, out, d_bad);
    else
        convert_index_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)dev, n, index_base, limit, out, d_bad);
    GNNB_LAUNCHED();
    if (staged) {
        GNNB_CUDA(cudaStreamSynchronize(st));
        cudaFree(staged);
    }
    return GNNB_OK;
}

}  // namespace gnnb

using namespace gnnb;

extern "C" {

int gnnb_set_chunk_edges(int chunk) {
    if (chunk < 32 || chunk > 4096 || (chunk & (chunk - 1))) GNNB_FAIL(GNNB_EINVAL, "chunk must be a power of two in [32,4096]");
    g_chunk_default = chunk;
    return GNNB_OK;
}

int gnnb_graph_create(gnnb_graph_t* out, const void* src, const void* dst, int64_t num_edges,
                      int64_t num_src, int64_t num_dst, int index_bytes, int index_base,
                      int on_device, void* stream) {
    if (!out) GNNB_FAIL(GNNB_EINVAL, "out handle is NULL");
    *out = nullptr;
    if (index_bytes != 4 && index_bytes != 8) GNNB_FAIL(GNNB_EINVAL, "index_bytes must be 4 or 8 (got %d)", index_bytes);
    if (index_base != 0 && index_base != 1) GNNB_FAIL(GNNB_EINVAL, "index_base must be 0 or 1 (got %d)", index_base);
    if (num_edges < 0 || num_src < 0 || num_dst < 0) GNNB_FAIL(GNNB_ESIZE, "negative size");
    if (num_edges >= ((int64_t)1 << 31) - 1 || num_src >= ((int64_t)1 << 31) - 1 || num_dst >= ((int64_t)1 << 31) - 1)
        GNNB_FAIL(GNNB_ESIZE, "a single plan is int32-indexed: E and N must be < 2^31-1 (partition larger graphs)");
    if (num_edges > 0 && (!src || !dst)) GNNB_FAIL(GNNB_EINVAL, "src/dst is NULL");
    if (gnnb_device_count() <= 0) GNNB_FAIL(GNNB_ECUDA, "no CUDA device: libgnnb200 has no CPU fallback");
    cudaStream_t st = (cudaStream_t)stream;
    gnnb_graph* g = new gnnb_graph();
    g->E = num_edges;
    g->n_src = (int32_t)num_src;
    g->n_dst = (int32_t)num_dst;
    g->chunk = g_chunk_default;
    cudaGetDevice(&g->device);
    int status = GNNB_OK;
    int* d_bad = nullptr;
    do {
#define GNNB_STEP(expr) { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); status = (_e == cudaErrorMemoryAllocation) ? GNNB_ENOMEM : GNNB_ECUDA; break; } }
        size_t nE = (size_t)(num_edges > 0 ? num_edges : 1);
        GNNB_STEP(cudaMalloc(&g->coo_src, sizeof(int32_t) * nE));
        GNNB_STEP(cudaMalloc(&g->coo_dst, sizeof(int32_t) * nE));
        GNNB_STEP(cudaMalloc(&d_bad, sizeof(int)));
        GNNB_STEP(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
        if ((status = convert_indices(src, num_edges, index_bytes, index_base, num_src, on_device, g->coo_src, d_bad, st))) break;
        if ((status = convert_indices(dst, num_edges, index_bytes, index_base, num_dst, on_device, g->coo_dst, d_bad, st))) break;
        int bad = 0;
        GNNB_STEP(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        GNNB_STEP(cudaStreamSynchronize(st));
        if (bad) {
            set_error("edge index out of range: every index must lie in [%d, num_nodes%s] (convert.jl:49-54)",
                      index_base, index_base ? "" : ")");
            status = GNNB_EINDEX;
            break;
        }
        if ((status = ensure_csr(g, false, st))) break;
#undef GNNB_STEP
    } while (0);
    cudaFree(d_bad);
    if (status != GNNB_OK) {
        gnnb_graph_destroy(g);
        return status;
    }
    *out = g;
    return GNNB_OK;
}

int gnnb_graph_destroy(gnnb_graph_t g) {
    if (!g) return GNNB_OK;
    cudaFree(g->coo_src);
    cudaFree(g->coo_dst);
    free_csr(g->by_dst);
    free_csr(g->by_src);
    cudaFree(g->ws);
    cudaFree(g->ws2);
    cudaFree(g->gcn_c);
    cudaFree(g->host_ws);
    delete g;
    return GNNB_OK;
}

int gnnb_graph_info(gnnb_graph_t g, int64_t* num_edges, int64_t* num_src, int64_t* num_dst) {
    if (!g) GNNB_FAIL(GNNB_EINVAL, "graph handle is NULL");
    if (num_edges) *num_edges = g->E;
    if (num_src) *num_src = g->n_src;
    if (num_dst) *num_dst = g->n_dst;
    return GNNB_OK;
}

static int derive_self_loop_csr(const Csr& o, Csr& c, int64_t E, int32_t n, int32_t chunk, cudaStream_t st) {
    GNNB_TRY(alloc_csr(c, E + n, n, n, chunk));
    if (E > 0) {
        selfloop_edges_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, st>>>(o.row, o.col, o.eid, E, c.row, c.col, c.eid);
        GNNB_LAUNCHED();
    }
    selfloop_rows_kernel<<<(unsigned)ceil_div((int64_t)n + 1, 256), 256, 0, st>>>(o.rowptr, n, (int32_t)E, c.rowptr,
                                                                              c.row, c.col, c.eid);
    GNNB_LAUNCHED();
    GNNB_TRY(find_long_rows(c, chunk, st));
    c.built = true;
    return GNNB_OK;
}

int gnnb_graph_add_self_loops(gnnb_graph_t g, gnnb_graph_t* out, void* stream) {
    if (!g || !out) GNNB_FAIL(GNNB_EINVAL, "NULL argument");
    *out = nullptr;
    if (g->n_src != g->n_dst) GNNB_FAIL(GNNB_ESIZE, "add_self_loops needs num_src == num_dst");
    const int32_t n = g->n_src;
    const int64_t E = g->E, E2 = E + n;
    if (E2 >= ((int64_t)1 << 31) - 1) GNNB_FAIL(GNNB_ESIZE, "E + N must be < 2^31-1");
    cudaStream_t st = (cudaStream_t)stream;
    gnnb_graph* h = new gnnb_graph();
    h->E = E2; h->n_src = n; h->n_dst = n; h->chunk = g->chunk; h->device = g->device;
    int status = GNNB_OK;
    do {
        size_t nE = (size_t)(E2 > 0 ? E2 : 1);
        if (cudaMalloc(&h->coo_src, sizeof(int32_t) * nE) != cudaSuccess ||
            cudaMalloc(&h->coo_dst, sizeof(int32_t) * nE) != cudaSuccess) {
            set_error("cudaMalloc failed in add_self_loops"); status = GNNB_ENOMEM; break;
        }
        if (E > 0) {
            cudaMemcpyAsync(h->coo_src, g->coo_src, sizeof(int32_t) * E, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(h->coo_dst, g->coo_dst, sizeof(int32_t) * E, cudaMemcpyDeviceToDevice, st);
        }
        if (n > 0) {
            iota_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(h->coo_src + E, n, 0);
            iota_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(h->coo_dst + E, n, 0);
            g_launches.fetch_add(2, std::memory_order_relaxed);
        }
        if ((status = ensure_csr(g, false, st))) break;
        if ((status = derive_self_loop_csr(g->by_dst, h->by_dst, E, n, h->chunk, st))) break;
        if (g->by_src.built) {
            if ((status = derive_self_loop_csr(g->by_src, h->by_src, E, n, h->chunk, st))) break;
        }
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("add_self_loops: %s", cudaGetErrorString(e)); status = GNNB_ECUDA; break; }
    } while (0);
    if (status != GNNB_OK) { gnnb_graph_destroy(h); return status; }
    *out = h;
    return GNNB_OK;
}

int gnnb_graph_csr(gnnb_graph_t g, int transposed, int32_t* rowptr, int32_t* col, int32_t* eid, void* stream) {
    if (!g) GNNB_FAIL(GNNB_EINVAL, "graph handle is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    GNNB_TRY(ensure_csr(g, transposed != 0, st));
    const Csr& c = transposed ? g->by_src : g->by_dst;
    if (rowptr) GNNB_CUDA(cudaMemcpyAsync(rowptr, c.rowptr, sizeof(int32_t) * ((size_t)c.nrows + 1), cudaMemcpyDeviceToHost, st));
    if (col && g->E) GNNB_CUDA(cudaMemcpyAsync(col, c.col, sizeof(int32_t) * (size_t)g->E, cudaMemcpyDeviceToHost, st));
    if (eid && g->E) GNNB_CUDA(cudaMemcpyAsync(eid, c.eid, sizeof(int32_t) * (size_t)g->E, cudaMemcpyDeviceToHost, st));
    GNNB_CUDA(cudaStreamSynchronize(st));
    return GNNB_OK;
}

}  // extern "C"
