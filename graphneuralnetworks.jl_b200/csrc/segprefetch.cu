// segprefetch.cu — A/B variants 6..9 of the fused segmented reduce for rows of 128 floats (one warp = one 512 B row).
//
// Same decomposition, same arithmetic in the same order as seg_reduce_kernel<4,32,1,·> (segreduce.cu) — results are
// bit-identical — with one change to the schedule: the per-edge index words of the NEXT group of 32 edges are requested
// around the current group's row reduction (col, row, w before it; the dependent cs[col] after it).  In the default kernel a warp starts every
// group with two dependent global loads (col, then cs[col]) during which it has no row load in flight; with 32 resident
// warps per SM and a kernel whose throughput follows the number of requests in flight (profiles/r1_seg_variants.md),
// those bubbles are about a fifth of a warp's time per chunk.  Variants 8, 9 additionally read the chunk decomposition
// (chunk_bounds) from a 16 B-per-chunk table built once per plan instead of recomputing it through three dependent
// loads in every warp's prologue.  Kept in its own translation unit so that the measured
// default kernels keep their exact code.  Written after round 1's GPU budget was spent: not yet measured
// (scripts/sweep_variants.py VARIANTS=0,6,7 checks bit-identity and times it).
#include "common.cuh"
#include "segwalk.cuh"
#include "segparams.cuh"
#include <math_constants.h>

namespace gnnb {

namespace {

template <bool ISMAX> __device__ __forceinline__ float pf_comb1(float acc, float v, float s1, float s2) {
    float m = __fmul_rn(__fmul_rn(v, s1), s2);          // (x * cs) * w, each product rounded, then the reduction
    if (ISMAX) return fmaxf(acc, m);
    return __fadd_rn(acc, m);
}
template <bool ISMAX> __device__ __forceinline__ float4 pf_comb(float4 a, float4 v, float s1, float s2) {
    return make_float4(pf_comb1<ISMAX>(a.x, v.x, s1, s2), pf_comb1<ISMAX>(a.y, v.y, s1, s2),
                       pf_comb1<ISMAX>(a.z, v.z, s1, s2), pf_comb1<ISMAX>(a.w, v.w, s1, s2));
}
__device__ __forceinline__ float4 pf_mul(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 pf_div(float4 a, float s) {
    return make_float4(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s), __fdiv_rn(a.w, s));
}
__device__ __forceinline__ float4 pf_finish(float4 acc, const SegParams& p, int r) {
    if (p.mean) {
        int deg = __ldg(p.rowptr + r + 1) - __ldg(p.rowptr + r);
        acc = pf_div(acc, (float)(deg > 0 ? deg : 1));
    }
    if (p.ct) acc = pf_mul(acc, __ldg(p.ct + r));
    if (p.sign < 0.f) acc = pf_mul(acc, -1.f);
    return acc;
}

// flags of a precomputed chunk record
constexpr int CI_HEAD = 1, CI_TAIL = 2;

// chunk_bounds() of every chunk, once per plan: the kernel prologue becomes one 16 B load instead of a chain of three
// dependent ones (row[a] -> rowptr[r0], rowptr[r0+1]; row[z-1] -> rowptr[r1] ...)
__global__ void chunk_info_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ row, int C, int E,
                                  int nchunks, int4* __restrict__ info) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    const ChunkBounds b = chunk_bounds(rowptr, row, k, C, E, nchunks);
    info[k] = make_int4(b.e_begin, b.e_end, b.prev_row, (b.head_partial ? CI_HEAD : 0) | (b.tail_partial ? CI_TAIL : 0));
}

template <bool ISMAX, int U>
__global__ void __launch_bounds__(256, 4) seg_reduce_prefetch_kernel(const SegParams p, const int32_t* __restrict__ chunk_info) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TPR = 32;
    const int lig = threadIdx.x % TPR;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / TPR;   // chunk id
    const int64_t foff = (int64_t)blockIdx.y * (4 * TPR) + (int64_t)lig * 4;    // this lane's 4 floats of the row
    const bool fact = foff < p.D;
    const float neutral = ISMAX ? -CUDART_INF_F : 0.f;
    const float fillv = ISMAX ? (p.sign < 0.f ? CUDART_INF_F : -CUDART_INF_F) : 0.f;

    int e_begin = 0, e_end = 0, prev_row = -1;
    bool head_partial = false, tail_partial = false;
    if (chunk_info != nullptr) {
        if (k < p.nchunks) {
            const int4 ci = __ldg(reinterpret_cast<const int4*>(chunk_info) + k);
            e_begin = ci.x; e_end = ci.y; prev_row = ci.z;
            head_partial = (ci.w & CI_HEAD) != 0;
            tail_partial = (ci.w & CI_TAIL) != 0;
        }
    } else {
        const ChunkBounds cb = chunk_bounds(p.rowptr, p.row, k, p.chunk, p.E, p.nchunks);
        e_begin = cb.e_begin; e_end = cb.e_end; prev_row = cb.prev_row;
        head_partial = cb.head_partial; tail_partial = cb.tail_partial;
    }

    float4 acc = make_float4(neutral, neutral, neutral, neutral);
    int r = -1;
    bool first_flush = true;

    auto flush = [&](int rr, bool last) {
        float* base;
        bool raw;
        if (first_flush && head_partial) {
            base = p.ws + (size_t)(2 * k + 0) * p.D;
            raw = true;
        } else if (last && tail_partial) {
            base = p.ws + (size_t)(2 * k + 1) * p.D;
            raw = true;
        } else {
            base = p.out + (size_t)rr * p.D;
            raw = false;
        }
        first_flush = false;
        if (fact) *reinterpret_cast<float4*>(base + foff) = raw ? acc : pf_finish(acc, p, rr);
    };
    auto fill_gap = [&](int lo, int hi) {
        if (!p.fill) return;
        for (int q = lo + 1; q < hi; ++q)
            if (fact) *reinterpret_cast<float4*>(p.out + (size_t)q * p.D + foff) = make_float4(fillv, fillv, fillv, fillv);
    };
    // Index words of the group of 32 edges starting at e0, one edge per lane, in two phases because a warp issues in
    // order: phase A (col, row, w) has no dependency and is requested BEFORE the current group's rows are reduced;
    // phase B (cs[col]) needs col and is requested AFTER them, when col has long arrived — requested any earlier it would
    // park the warp for a full memory latency with no row load in flight, which is exactly the bubble to remove.
    auto load_a = [&](int e0, int& c, int& d, float& s2) {
        const int pe = e0 + lig;
        c = 0; d = -1; s2 = 1.f;
        if (pe < e_end) {
            c = __ldg(p.col + pe);
            d = __ldg(p.row + pe);
            if (p.w) s2 = __ldg(p.w + pe);
            if (p.sign < 0.f) s2 = -s2;
        }
    };
    auto load_b = [&](int e0, int c) -> float {
        return (p.cs != nullptr && e0 + lig < e_end) ? __ldg(p.cs + c) : 1.f;
    };

    int e = e_begin;
    int c_n, d_n;
    float s1_n, s2_n;
    load_a(e, c_n, d_n, s2_n);
    s1_n = load_b(e, c_n);
    while (__any_sync(FULL, e < e_end)) {
        const int c_l = c_n, d_l = d_n;
        const float s1_l = s1_n, s2_l = s2_n;
        load_a(e + TPR, c_n, d_n, s2_n);                      // in flight while this group's rows are reduced
        const int nb = e_end - e;
#pragma unroll 1
        for (int j0 = 0; j0 < TPR; j0 += U) {
            if (!__any_sync(FULL, j0 < nb)) break;
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int cj = __shfl_sync(FULL, c_l, (j0 + u) & 31);
                const bool valid = (j0 + u) < nb && (j0 + u) < TPR;
                const float* xr = (p.x2 != nullptr && cj >= p.split) ? p.x2 + (size_t)(cj - p.split) * p.D
                                                                     : p.x + (size_t)cj * p.D;
                v[u] = (valid && fact) ? __ldg(reinterpret_cast<const float4*>(xr + foff)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int dj = __shfl_sync(FULL, d_l, (j0 + u) & 31);
                const float s1 = __shfl_sync(FULL, s1_l, (j0 + u) & 31);
                const float s2 = __shfl_sync(FULL, s2_l, (j0 + u) & 31);
                if ((j0 + u) < nb && (j0 + u) < TPR) {
                    if (dj != r) {
                        if (r >= 0) {
                            flush(r, false);
                            fill_gap(r, dj);
                        } else if (!head_partial) {
                            fill_gap(prev_row, dj);
                        }
                        r = dj;
                        acc = make_float4(neutral, neutral, neutral, neutral);
                    }
                    acc = pf_comb<ISMAX>(acc, v[u], s1, s2);
                }
            }
        }
        s1_n = load_b(e + TPR, c_n);                          // col of the next group is here by now: no stall
        e += TPR;
    }
    if (e_begin < e_end) {
        flush(r, true);
        if (e_end == p.E) fill_gap(r, p.nrows);
    }
}

}  // namespace

// variants 8, 9: build (once per plan and direction) and attach the precomputed chunk records
int ensure_chunk_info(gnnb_graph* g, const Csr& c, const int32_t** out, cudaStream_t st) {
    Csr& mc = const_cast<Csr&>(c);          // c is g->by_dst or g->by_src, both owned (mutably) by the plan
    const int32_t nchunks = (int32_t)ceil_div(g->E, g->chunk);
    if (mc.chunk_info == nullptr && nchunks > 0) {
        std::lock_guard<std::mutex> lock(g->mu);
        if (mc.chunk_info == nullptr) {
            int32_t* buf = nullptr;
            GNNB_CUDA(cudaMalloc(&buf, sizeof(int4) * (size_t)nchunks));
            chunk_info_kernel<<<(unsigned)ceil_div(nchunks, 256), 256, 0, st>>>(c.rowptr, c.row, g->chunk, (int)g->E, nchunks,
                                                                             reinterpret_cast<int4*>(buf));
            GNNB_LAUNCHED();
            mc.chunk_info = buf;
        }
    }
    *out = mc.chunk_info;
    return GNNB_OK;
}

int launch_seg_prefetch(const SegParams& p, const int32_t* chunk_info, bool ismax, int loads, dim3 grid, cudaStream_t st) {
    if (loads == 8) {
        if (ismax) seg_reduce_prefetch_kernel<true, 8><<<grid, 256, 0, st>>>(p, chunk_info);
        else seg_reduce_prefetch_kernel<false, 8><<<grid, 256, 0, st>>>(p, chunk_info);
    } else {
        if (ismax) seg_reduce_prefetch_kernel<true, 6><<<grid, 256, 0, st>>>(p, chunk_info);
        else seg_reduce_prefetch_kernel<false, 6><<<grid, 256, 0, st>>>(p, chunk_info);
    }
    GNNB_LAUNCHED();
    return GNNB_OK;
}

}  // namespace gnnb
