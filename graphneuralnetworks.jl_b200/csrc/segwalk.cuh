// segwalk.cuh — the chunk decomposition shared by every segmented kernel (segreduce.cu, gat.cu).
//
// The CSR-sorted edge list is cut into chunks of C edges; chunk k is owned by one group of lanes.
//   * a row that starts inside chunk k and has <= C edges is finished by chunk k (possibly overrunning
//     into chunk k+1, which skips those edges);
//   * a row with > C edges is "long": every chunk it touches reduces its own piece and stores a partial
//     (slot 2k+1 for the piece in the chunk where the row starts, slot 2k for every later chunk); a fix-up
//     kernel combines the slots in chunk order.
#pragma once
#include <stdint.h>

namespace gnnb {

struct ChunkBounds {
    int e_begin, e_end;   // edges this group processes
    bool head_partial;    // first processed row is a piece of a long row begun earlier -> slot 2k
    bool tail_partial;    // last processed row is a long row continuing past the chunk -> slot 2k+1
    int prev_row;         // last non-empty row before e_begin (-1 if none); valid when !head_partial
};

__device__ __forceinline__ ChunkBounds chunk_bounds(const int32_t* __restrict__ rowptr,
                                                    const int32_t* __restrict__ row, int64_t k, int C,
                                                    int E, int nchunks) {
    ChunkBounds b;
    b.e_begin = 0; b.e_end = 0; b.head_partial = false; b.tail_partial = false; b.prev_row = -1;
    if (k >= nchunks) return b;
    const int a = (int)(k * C);
    const int z = (a + C < E) ? a + C : E;
    b.e_begin = a;
    b.e_end = z;
    const int r0 = __ldg(row + a);
    const int rs0 = __ldg(rowptr + r0), re0 = __ldg(rowptr + r0 + 1);
    if (rs0 < a) {                       // row r0 began in an earlier chunk
        if (re0 - rs0 > C) b.head_partial = true;   // long row: we own the piece [a, ..)
        else b.e_begin = re0;                        // short row: its first chunk finishes it
    }
    if (b.e_begin < z) {
        const int r1 = __ldg(row + z - 1);
        const int rs1 = __ldg(rowptr + r1), re1 = __ldg(rowptr + r1 + 1);
        if (re1 > z) {                   // last row continues past the chunk
            if (re1 - rs1 > C) b.tail_partial = true;  // long: piece [.., z)
            else b.e_end = re1;                         // short: overrun and finish it
        }
        if (!b.head_partial) b.prev_row = (b.e_begin > 0) ? __ldg(row + b.e_begin - 1) : -1;
    } else {
        b.e_end = b.e_begin;             // nothing left for this chunk
    }
    return b;
}

}  // namespace gnnb
