// segparams.cuh — launch parameters shared by the register-staged segmented-reduce kernels (segreduce.cu, segprefetch.cu)
#pragma once
#include <stdint.h>

namespace gnnb {

struct SegParams {
    const int32_t* __restrict__ rowptr;
    const int32_t* __restrict__ col;
    const int32_t* __restrict__ row;
    const float* __restrict__ x;
    const float* __restrict__ x2;   // rows of gathered nodes >= split live here (halo buffer); nullptr = single base
    const float* __restrict__ w;
    const float* __restrict__ cs;
    const float* __restrict__ ct;
    float* __restrict__ out;
    float* __restrict__ ws;
    int64_t D;      // features per row (row stride)
    int32_t E;
    int32_t nrows;
    int32_t chunk;
    int32_t nchunks;
    int32_t mean;   // divide by the row's edge count at the final store
    int32_t fill;   // 1: groups write the neutral element into the empty rows they pass over
    int32_t split;  // first gathered-node id served from x2
    float sign;     // +1, or -1 to turn MAX into MIN (min(m) = -max(-m))
};

}  // namespace gnnb
