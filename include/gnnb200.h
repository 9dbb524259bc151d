/*
 * gnnb200.h — C ABI of libgnnb200.so, the B200-native (sm_100a) message-passing engine that sits
 * behind GNNlib.jl's `propagate` / `apply_edges` / `aggregate_neighbors` hot path.
 *
 * Every entry point below replaces one reference interface; the citation after "replaces:" is
 * file:line under the reference tree (CarloLucibello/GraphNeuralNetworks.jl @ e46d1b04).  The Julia
 * binding a maintainer would add (`ccall`s from a package extension that mirrors
 * GNNlib/ext/GNNlibCUDAExt.jl:13-32) is shown in INTEGRATION.md and julia/GNNlibB200Ext.jl.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / CUDA types (`stream` is a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).
 *   - every function returns a gnnb_status; 0 = ok.  gnnb_last_error() gives the thread-local text.
 *     The host shim maps GNNB_ESIZE -> AssertionError (GNNGraphs/src/utils.jl:1-28) and
 *     GNNB_EINVAL -> ArgumentError (GNNlib/src/layers/conv.jl:3-10,22).
 *   - feature arrays are Julia column-major (D, N): node n owns D contiguous floats at x + n*D.
 *     Edge arrays (K, E) likewise, in the COO order of the graph the handle was created from.
 *   - "device" entry points take device pointers, are asynchronous on `stream`, never allocate
 *     caller-visible memory and never free caller memory.  The only library-owned object is the
 *     opaque graph plan.  "_host" entry points take HOST pointers and do the H2D/D2H themselves.
 *   - there is no CPU fallback: without a CUDA device every compute entry returns GNNB_ECUDA.
 */
#ifndef GNNB200_H
#define GNNB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GNNB_OK = 0,
    GNNB_EINVAL = 1,       /* bad argument value        -> ArgumentError  */
    GNNB_ESIZE = 2,        /* size mismatch             -> AssertionError */
    GNNB_ECUDA = 3,        /* CUDA runtime error / no device              */
    GNNB_ENOMEM = 4,       /* device allocation failed                    */
    GNNB_EUNSUPPORTED = 5, /* valid in the reference, not in this build   */
    GNNB_EINDEX = 6        /* node index out of [base, base+N) -> AssertionError (convert.jl:49-54) */
} gnnb_status;

/* message functions with a fused path (GNNlib/src/msgpass.jl:162-208) */
typedef enum {
    GNNB_COPY_XJ = 0,  /* copy_xj(xi,xj,e) = xj                 msgpass.jl:165 */
    GNNB_W_MUL_XJ = 1  /* w_mul_xj / e_mul_xj with a vector e   msgpass.jl:191-208 */
} gnnb_msg;

/* aggregations NNlib.scatter supports that the layers use */
typedef enum { GNNB_SUM = 0, GNNB_MEAN = 1, GNNB_MAX = 2, GNNB_MIN = 3 } gnnb_aggr;

/* which end of an edge */
typedef enum { GNNB_SRC = 0, GNNB_DST = 1 } gnnb_end;

/* degree direction (GNNGraphs/src/query.jl:314-369 `dir`) */
typedef enum { GNNB_DIR_OUT = 0, GNNB_DIR_IN = 1, GNNB_DIR_BOTH = 2 } gnnb_dir;

typedef struct gnnb_graph* gnnb_graph_t;

/* ---------------------------------------------------------------- library */

/* thread-local text of the last failure on this thread ("" if none) */
const char* gnnb_last_error(void);
/* library version string, e.g. "gnnb200 0.1 sm_100a" */
const char* gnnb_version(void);
/* number of CUDA devices visible (0 when there is none; never fails) */
int gnnb_device_count(void);
/* kernels launched by this library on the calling process since load (bench.py `gpu_launches`) */
int64_t gnnb_launch_count(void);

/* ------------------------------------------------------------------ graph
 * replaces: the COO GNNGraph value `(s, t)` that `edge_index(g)` hands to the hot path
 *           (GNNGraphs/src/query.jl:12-14, GNNGraphs/src/gnngraph.jl:108-117) plus the range
 *           validation of `to_coo` (GNNGraphs/src/convert.jl:49-54).
 * Builds the device plan: int32 0-based COO copy, CSR-by-target (stable in COO order: rowptr, col,
 * sorted targets, edge-id permutation) and, lazily on first backward/transposed use, CSR-by-source.
 *   src,dst      : E indices, `index_bytes` = 4 (Int32) or 8 (Int64), `index_base` = 1 (Julia) or 0
 *   num_src/dst  : node counts of the two ends (equal for a GNNGraph; they differ for the
 *                  [local|halo] source space of a node-partitioned shard)
 *   on_device    : 0 = src/dst are host pointers, 1 = device pointers
 * Errors: GNNB_EINDEX if any index is outside [base, base+num); GNNB_ESIZE if E or N is negative or
 *         E >= 2^31 (per-GPU shards are int32-indexed), GNNB_EINVAL for bad index_bytes/base. */
int gnnb_graph_create(gnnb_graph_t* out, const void* src, const void* dst, int64_t num_edges,
                      int64_t num_src, int64_t num_dst, int index_bytes, int index_base,
                      int on_device, void* stream);
int gnnb_graph_destroy(gnnb_graph_t g);

/* replaces: add_self_loops(g::GNNGraph{<:COO_T}) (GNNGraphs/src/transform.jl:12-28): a NEW plan
 * whose COO is [s; 1:n], [t; 1:n] (loops appended after the originals, existing loops kept).
 * Requires num_src == num_dst (GNNB_ESIZE otherwise). */
int gnnb_graph_add_self_loops(gnnb_graph_t g, gnnb_graph_t* out, void* stream);

/* num_edges, num_src, num_dst of a plan */
int gnnb_graph_info(gnnb_graph_t g, int64_t* num_edges, int64_t* num_src, int64_t* num_dst);

/* Copy the plan's index arrays to HOST buffers (NULL = skip) for bit-exact checks:
 *   transposed=0: CSR by target  (rowptr[num_dst+1], col = source of each sorted edge)
 *   transposed=1: CSR by source  (rowptr[num_src+1], col = target of each sorted edge)
 *   eid[k] = 0-based COO position of sorted edge k (stable: ascending within a row).
 * replaces nothing in the reference (it has no CSR type); rowptr differences must
 * equal degree(g; dir=:in / :out) exactly (GNNGraphs/test/query.jl:49-58). */
int gnnb_graph_csr(gnnb_graph_t g, int transposed, int32_t* rowptr, int32_t* col, int32_t* eid,
                   void* stream);

/* replaces: degree(g, Float32; dir, edge_weight) -> _degree (GNNGraphs/src/query.jl:314-331,355-369)
 * w = NULL -> counts (exact integers stored as float), else sum of w (COO order, length E).
 * out has num_dst (IN), num_src (OUT) entries; BOTH requires num_src == num_dst. */
int gnnb_degree(gnnb_graph_t g, int dir, const float* w, float* out, void* stream);

/* ------------------------------------------------------- gather / scatter
 * replaces: GNNGraphs._gather -> NNlib.gather (GNNGraphs/src/gatherscatter.jl:1-5):
 *           out[:,k] = x[:, idx[k]], idx = s (GNNB_SRC) or t (GNNB_DST), COO order. (D,N)->(D,E) */
int gnnb_gather(gnnb_graph_t g, int which, const float* x, int64_t D, float* out, void* stream);

/* replaces: GNNGraphs._scatter -> NNlib.scatter(aggr, m, t; dstsize=(D,n))
 *           (GNNGraphs/src/gatherscatter.jl:12-18; called from aggregate_neighbors, msgpass.jl:145-149)
 * m is (D,E) in COO order, out (D, num_dst).  Empty targets get the op's neutral element:
 * 0 for SUM/MEAN, -Inf for MAX, +Inf for MIN (NNlib semantics).
 * which = GNNB_DST scatters by target (aggregate_neighbors); GNNB_SRC scatters by source into
 * (D, num_src) (the pullback of gather by s). */
int gnnb_scatter(gnnb_graph_t g, int which, int aggr, const float* m, int64_t D, float* out,
                 void* stream);

/* ------------------------------------------------------------- propagate
 * replaces: propagate(copy_xj | w_mul_xj | e_mul_xj(vector e), g, + | mean | max | min; xj)
 *           = aggregate_neighbors(g, aggr, apply_edges(f, g, xi, xj, e))   (msgpass.jl:71-79),
 *           its CPU SpMM specialisations (msgpass.jl:215-238) and the CUDA-ext re-routing
 *           (GNNlib/ext/GNNlibCUDAExt.jl:13-32) — fused: no (D,E) intermediate.
 *   out[:,i] = ct[i] * AGG_{k in N(i)} ( w[k] * cs[s_k] * x[:, s_k] )
 *   x  (D, num_src), out (D, num_dst); w NULL or E floats in COO order (required for W_MUL_XJ);
 *   cs NULL or num_src floats, ct NULL or num_dst floats: optional per-node scales fused into the
 *   load / store (GCN's 1/sqrt(d), conv.jl:57-67).  MEAN divides by the in-degree (0 for isolated).
 *   transposed != 0 runs the same reduction on the reversed graph (x is (D,num_dst), out (D,num_src)):
 *   that is the pullback of the SUM/MEAN forward w.r.t. xj. */
int gnnb_propagate(gnnb_graph_t g, int transposed, int msg, int aggr, const float* x,
                   const float* w, const float* cs, const float* ct, int64_t D, float* out,
                   void* stream);

/* Pullbacks of gnnb_propagate (what Zygote computes today through NNlib's rrules).
 *   dout (D,num_dst).  dx (D,num_src) or NULL.  dw (E, COO order) or NULL (W_MUL_XJ only).
 *   SUM/MEAN: dx[:,j] = cs[j] * sum_{k: s_k=j} w_k * ct'[t_k] * dout[:,t_k]   (ct' = ct/deg for MEAN)
 *             dw[k]   = ct'[t_k] * cs[s_k] * <dout[:,t_k], x[:,s_k]>
 *   MAX/MIN : needs x and the forward output `out_fwd`; every tied arg-extremum receives the
 *             gradient (NNlib rule): dx[:,j] += w_k cs[j] ct[t_k] dout[:,t_k] .* (m_k .== out_fwd[:,t_k]/ct[t_k])
 *             (dw unsupported for MAX/MIN: GNNB_EUNSUPPORTED). */
int gnnb_propagate_bwd(gnnb_graph_t g, int msg, int aggr, const float* dout, const float* x,
                       const float* w, const float* cs, const float* ct, const float* out_fwd,
                       int64_t D, float* dx, float* dw, void* stream);

/* ------------------------------------------------------------ edge softmax
 * replaces: softmax_edge_neighbors(g, e) (GNNlib/src/utils.jl:84-97): for every leading index,
 *           a softmax over the edges that share a target.  e, out are (K,E) in COO order. */
int gnnb_softmax_edge_neighbors(gnnb_graph_t g, const float* e, int64_t K, float* out, void* stream);
/* pullback: de_k = a_k (da_k - sum_{k' in N(i)} a_k' da_k'), a = forward output. */
int gnnb_softmax_edge_neighbors_bwd(gnnb_graph_t g, const float* alpha, const float* dalpha,
                                    int64_t K, float* de, void* stream);

/* --------------------------------------------------------------- GCN core
 * replaces: the message-passing core of gcn_conv (GNNlib/src/layers/conv.jl:52-67):
 *     d = degree(g, T; dir=:in, edge_weight); c = 1 ./ sqrt.(d)       (default norm_fn, conv.jl:99)
 *     out = (propagate(copy_xj | e_mul_xj, g, +, xj = x .* c')) .* c'
 * `g` must already carry the self loops if the layer adds them (gnnb_graph_add_self_loops).
 * w NULL or E floats (COO order of g, loop weights included).  c_out (num_dst floats) receives c
 * (kept by the caller for the backward).  transposed=1 computes the pullback w.r.t. x given dout
 * and the forward's c:  dx = c .* (A^T-propagate(dout .* c)).
 * c == NULL (w must be NULL too): the plan's own c = 1 ./ sqrt.(in-degree), computed once and kept with the plan
 * together with its per-edge stream c[s_k] in plan order, which spares the kernel one dependent gather per edge. */
int gnnb_gcn_norm(gnnb_graph_t g, const float* w, float* c_out, void* stream);
int gnnb_gcn_propagate(gnnb_graph_t g, int transposed, const float* x, const float* w,
                       const float* c, int64_t D, float* out, void* stream);

/* --------------------------------------------------------------- GAT core
 * replaces: the edge part of gat_conv + gat_message (GNNlib/src/layers/conv.jl:136-141,152-167):
 *     logα_k[h] = leakyrelu(el[h,t_k] + er[h,s_k], slope)          (a·[Wx_i;Wx_j])
 *     α = softmax_edge_neighbors(g, logα);  out[:,h,i] = Σ_k α_k[h] Wx[:,h,s_k]
 * Wx (C,H,num_src), el (H,num_dst), er (H,num_src), out (C,H,num_dst).
 * Optional outputs (NULL = skip): alpha (H,E) in COO order; seg_max, seg_sum (H,num_dst) — the
 * per-target softmax statistics the backward recomputes α from. */
int gnnb_gat_aggregate(gnnb_graph_t g, const float* Wx, const float* el, const float* er,
                       int64_t C, int64_t H, float slope, float* out, float* alpha,
                       float* seg_max, float* seg_sum, void* stream);
/* pullback: given dout (C,H,num_dst), the forward output and the forward statistics, produce
 *   dWx (C,H,num_src) = Σ_{k: s_k=j} α_k dout[:,:,t_k]        (attention-weighted transposed pull, α recomputed)
 *   del (H,num_dst), der (H,num_src): gradients of the two per-node logit terms
 * (the el/er -> a, Wx chain is dense per-node work left to the caller's AD). */
int gnnb_gat_aggregate_bwd(gnnb_graph_t g, const float* Wx, const float* el, const float* er,
                           const float* seg_max, const float* seg_sum, const float* out_fwd,
                           const float* dout, int64_t C, int64_t H, float slope, float* dWx, float* del,
                           float* der, void* stream);

/* The per-node halves of the attention logits (csrc/gatlogit.cu) — replaces `sum(l.a .* vcat(Wxi, Wxj), dims = 1)` of
 * gat_message (GNNlib/src/layers/conv.jl:157-163), which splits into a target and a source term:
 *     el[h,i] = sum_c a[c,h] Wx[c,h,i]          er[h,j] = sum_c a[C+c,h] Wx[c,h,j]
 * Wx (C,H,N), a (2C,H) column-major as the layer stores it, el / er (H,N).  One pass over Wx.
 * Pullback: dWx_accum (C,H,N) += del[h,n] a[c,h] + der[h,n] a[C+c,h]  IN PLACE (it already holds the dWx of
 * gnnb_gat_aggregate_bwd), da (2C,H) = [sum_n del Wx ; sum_n der Wx] — deterministic (fixed-order reduction).
 * Shapes: C/4 a power of two <= 32, C*H <= 4096 (GNNB_EUNSUPPORTED otherwise). */
int gnnb_gat_logit_terms(const float* Wx, const float* a, int64_t N, int64_t C, int64_t H, float* el, float* er, void* stream);
int gnnb_gat_logit_terms_bwd(const float* Wx, const float* a, const float* del, const float* der, int64_t N, int64_t C,
                             int64_t H, float* dWx_accum, float* da, void* stream);

/* ------------------------------------------------------- dense layer part
 * replaces: l.σ.(weight * x .+ l.bias) of the conv layers (GNNlib/src/layers/conv.jl:39,69-71; :281) and its pullback.
 * Din, Dout <= 128: hand-written tcgen05 kernels (csrc/dense_tc.cu: 3xTF32 split, TMEM accumulators, bias/relu epilogue
 * written as whole row segments, W resident in shared memory); Din % 32 == 0 <= 2048 and Dout % 128 == 0 <= 1024 with at
 * least 2048 nodes: the wide tcgen05 kernel of the same file (both operands streamed, W through cp.async.bulk; forward and
 * dx); every other shape: a library GEMM like the reference's, issued through cuBLASLt 12.9
 * with the fp32-emulated compute type (bf16 x9, fp32 accumulate; SIMT sgemm if unavailable), bias (+relu) in the epilogue.  x (Din,N), W (Dout,Din) row-major as the layer stores it, bias NULL or Dout floats, y (Dout,N).
 * relu: 0 = identity, 1 = relu. */
int gnnb_linear(const float* x, const float* W, const float* bias, int relu, int64_t N, int64_t Din,
                int64_t Dout, float* y, void* stream);
/* pullback: dy (Dout,N); y = forward output (relu only); dpre_ws = (Dout,N) workspace (relu only);
 * outputs (each may be NULL): dx (Din,N), dW (Dout,Din), db (Dout).  The relu mask x upstream gradient and the bias
 * gradient are one hand-written pass (deterministic two-stage column sum). */
int gnnb_linear_bwd(const float* dy, const float* y, const float* x, const float* W, int relu, int64_t N,
                    int64_t Din, int64_t Dout, float* dpre_ws, float* dx, float* dW, float* db, void* stream);
/* σ.(W * vcat(x1, x2) .+ b) — sage_conv's dense part (GNNlib/src/layers/conv.jl:281) — without the (Din1+Din2, N) vcat
 * temporary: the two column blocks of W (Dout, Din1+Din2, row-major as the layer stores it) meet x1 (Din1,N) and x2 (Din2,N)
 * in two passes of the tcgen05 kernel, the second adding the first's result before bias / activation.  Pullback: dx1, dx2,
 * dW (Dout, Din1+Din2), db, each may be NULL.  Shapes: Din1, Din2 multiples of 32 <= 128, Dout = 128 (forward also Dout a
 * multiple of 16 <= 128); GNNB_EUNSUPPORTED otherwise (callers concatenate and use gnnb_linear). */
int gnnb_linear2(const float* x1, const float* x2, const float* W, const float* bias, int relu, int64_t N, int64_t Din1,
                 int64_t Din2, int64_t Dout, float* y, void* stream);
int gnnb_linear2_bwd(const float* dy, const float* y, const float* x1, const float* x2, const float* W, int relu, int64_t N,
                     int64_t Din1, int64_t Din2, int64_t Dout, float* dpre_ws, float* dx1, float* dx2, float* dW, float* db,
                     void* stream);
/* y = act(x .+ bias) on (D,N) features and its pullback, for layers whose closing `σ.(x .+ bias)` follows an aggregation
 * rather than a GEMM (GATConv: GNNlib/src/layers/conv.jl:149): one pass each instead of the broadcast-add, the
 * activation, the mask product and the column reduction.  bias NULL or D floats; relu 0/1; y may alias x.
 * bwd: dpre = dy .* (y > 0) (relu only; dpre may alias dy), db = sum over nodes of dpre (NULL to skip; deterministic). */
int gnnb_bias_act(const float* x, const float* bias, int relu, int64_t N, int64_t D, float* y, void* stream);
int gnnb_bias_act_bwd(const float* dy, const float* y, int relu, int64_t N, int64_t D, float* dpre, float* db, void* stream);
/* 1 (default) = try the fp32-emulated tensor-core GEMM; 0 = force the SIMT sgemm.  *_active: -1 not yet used,
 * 0 unavailable / off, 1 in use. */
int gnnb_dense_set_emulation(int on);
/* The hand-written tcgen05 kernels (csrc/dense_tc.cu: 3xTF32 split, TMEM accumulators, bias/relu epilogue) serve
 * Din, Dout <= 128 (Din % 32 == 0, Dout % 16 == 0) and the wide shapes named above for gnnb_linear and the dx part of
 * gnnb_linear_bwd (dW: Dout == 128 only); 0 switches them off (cuBLASLt everywhere).  gnnb_dense_tc_error() != 0 means one of its bounded pipeline waits expired. */
int gnnb_dense_set_tensor_core_kernel(int on);
int gnnb_dense_tc_error(void);
int gnnb_dense_emulation_active(void);

/* ------------------------------------------------- node-partitioned shards
 * (no reference counterpart: the reference has no distributed code)
 * A shard is an ordinary plan whose targets are the nodes one GPU owns (num_dst = n_local) and whose
 * source ids live in the space [local rows | halo rows] (num_src = n_local + n_halo): created with
 * gnnb_graph_create(..., num_src, num_dst, ...).  The halo rows arrive through one all-to-all-v per pass
 * (NCCL, driven by the host side: graphneuralnetworks.jl_b200/partition.py).
 *
 * gnnb_gather_rows: out[k,:] = x[idx[k],:] for an explicit int32 0-based DEVICE index list — packs the rows
 * a peer requested into the send buffer. */
int gnnb_gather_rows(const int32_t* idx_dev, int64_t n, const float* x, int64_t D, float* out,
                     void* stream);
/* gnnb_propagate_halo: gnnb_propagate (forward direction of the shard plan) with the gathered rows split
 * over two buffers: node ids < n_local read x_local, the others read x_halo + (id - n_local)*D.
 * cs (optional) has num_src entries ([local | halo] order), ct num_dst. */
int gnnb_propagate_halo(gnnb_graph_t g, int msg, int aggr, const float* x_local, const float* x_halo,
                        int64_t n_local, const float* w, const float* cs, const float* ct, int64_t D,
                        float* out, void* stream);

/* Building one rank's shards on its GPU from chunks of the global edge list (csrc/shard.cu).
 *   ownership mode 0: contiguous ranges, bounds_host[q] <= v < bounds_host[q+1] (world + 1 entries; NULL = equal ranges);
 *   mode 1: cyclic — node v (0-based) belongs to rank v % world as its local row v / world, which spreads the hubs of a
 *   skewed id space over all ranks.  relabel_dev (mode 1, optional, DEVICE int32[num_nodes], alive until destroy):
 *   a permutation node -> position; the cyclic rule is applied to the position.  With positions = rank of the node by
 *   decreasing degree, the nodes are dealt to the ranks like cards: edge counts, node counts and the rows every rank must
 *   serve to its peers are all balanced, whatever the id space looks like (RMAT probabilities are products over id
 *   bits, so neither id ranges nor id % world balance it).
 * add(): a chunk of the global COO (DEVICE arrays, index_bytes 4|8, index_base 0|1; GNNB_EINDEX if out of range); keeps,
 *   stably, the edges whose target this rank owns (forward shard, direction 0) and those whose source it owns (backward
 *   shard, direction 1).  Synchronises the stream.
 * finish(direction): deduplicated sorted halo list, gathered nodes renamed into [local | halo], optional self loops
 *   (i, i) appended after the originals, plan created (*plan_out: an ordinary gnnb_graph_t, num_dst = n_local,
 *   num_src = n_local + n_halo; the caller destroys it).  recv_counts_host[q] (world entries) = halo rows owned by rank q.
 * halo(direction): the owner-local row index (int32, 0-based) of every halo entry, grouped by owner in rank order — what
 *   this rank asks each owner to send (n_halo entries, DEVICE). */
typedef struct gnnb_shard_builder* gnnb_shard_builder_t;
int gnnb_shard_builder_create(gnnb_shard_builder_t* out, int64_t num_nodes, int world, int rank, int mode,
                              const int64_t* bounds_host, const int32_t* relabel_dev);
int gnnb_shard_builder_add(gnnb_shard_builder_t b, const void* src, const void* dst, int64_t n, int index_bytes,
                           int index_base, void* stream);
int gnnb_shard_builder_finish(gnnb_shard_builder_t b, int direction, int add_self_loops, gnnb_graph_t* plan_out,
                              int64_t* n_local_out, int64_t* n_halo_out, int64_t* num_edges_out, int64_t* recv_counts_host,
                              void* stream);
int gnnb_shard_builder_halo(gnnb_shard_builder_t b, int direction, int32_t* halo_local_dev, void* stream);
int gnnb_shard_builder_destroy(gnnb_shard_builder_t b);
/* The relabel table of the degree-balanced deal: cost_dev (int32[num_nodes], zeroed by the caller) accumulates in + out
 * degree over chunks of the edge list; gnnb_balanced_relabel sorts the nodes by decreasing cost (stable) and deals them to
 * the ranks in turn, reversing direction every round: relabel_dev[node] = position (what gnnb_shard_builder_create takes),
 * order_dev[position] = node (rank q's local row k is node order_dev[k * world + q]). */
int gnnb_degree_accumulate(const void* src, const void* dst, int64_t n, int index_bytes, int index_base, int64_t num_nodes,
                           int32_t* cost_dev, void* stream);
int gnnb_balanced_relabel(const int32_t* cost_dev, int64_t num_nodes, int world, int32_t* relabel_dev, int32_t* order_dev,
                          void* stream);

/* Halo exchange without a staging copy: every rank writes the rows a peer asked for straight into that peer's halo
 * buffer over NVLink (peer-mapped memory).  gnnb_dev_alloc / gnnb_ipc_* manage exportable device buffers (plain
 * cudaMalloc + CUDA IPC handles, 64 bytes each, exchanged by the host side); gnnb_halo_push launches ONE kernel:
 * row k of the send list (grouped by peer: rows [seg_start[p], seg_start[p+1]) go to peer p) is copied from
 * x[send_idx[k],:] to peer_base[p] + (peer_row0[p] + k - seg_start[p])*D.  seg_start (world+1), peer_base (world),
 * peer_row0 (world) are HOST arrays.  Completion on the peers is published by a collective the caller issues on the
 * same stream afterwards. */
int gnnb_dev_alloc(void** p, int64_t bytes);
int gnnb_dev_free(void* p);
int gnnb_ipc_get_handle(void* p, unsigned char* handle64);
int gnnb_ipc_open_handle(const unsigned char* handle64, void** p);
int gnnb_ipc_close_handle(void* p);
int gnnb_halo_push(const int32_t* send_idx_dev, const int64_t* seg_start_host, const void* const* peer_base_host,
                   const int64_t* peer_row0_host, int world, const float* x, int64_t D, void* stream);

/* ------------------------------------------------- edge-list transforms
 * replaces: sort_edge_index(u, v) (GNNGraphs/src/utils.jl:41-45: sortperm of the zipped pairs, lexicographic and
 *           stable) — for CuArrays the reference's CUDA extension copies both arrays to the host, sorts there and copies
 *           back (GNNGraphs/ext/GNNGraphsCUDAExt.jl:24-30).  Here: one 64-bit key per pair, stable device radix sort.
 * u, v, u_out, v_out: E DEVICE integers of `index_bytes` (4|8) with values in [0, max_index] (0- or 1-based ids alike;
 * GNNB_EINDEX otherwise); outputs may alias the inputs or be NULL.  perm_out (NULL or E int64, 0-based): sorted
 * position k holds input pair perm_out[k].  Synchronises the stream. */
int gnnb_sort_edge_index(const void* u, const void* v, int64_t num_edges, int64_t max_index, int index_bytes,
                         void* u_out, void* v_out, int64_t* perm_out, void* stream);
/* replaces: the index half of remove_multi_edges(g; aggr) (GNNGraphs/src/transform.jl:157-190: edge_encoding, sortperm,
 *           first-occurrence mask, running segment id) and with it to_bidirected (transform.jl:495-510).
 * src, dst: E DEVICE indices in [index_base, index_base + num_nodes) (GNNB_EINDEX otherwise).  Outputs (DEVICE,
 * caller-allocated with E entries each): src_out/dst_out — the distinct pairs in (src, dst) order, first *num_unique
 * entries valid, same width and base as the input; perm_out (int64, 0-based) — the stable sort permutation; seg_out
 * (int64, 1-based) — which distinct pair sorted edge k collapsed into, i.e. the `idxs` the reference hands to
 * `_scatter(aggr, w[perm], idxs)`; that scatter is gnnb_scatter on a plan built from (1:E, seg_out).
 * num_unique is a HOST int64.  Synchronises the stream. */
int gnnb_coalesce_edges(const void* src, const void* dst, int64_t num_edges, int64_t num_nodes, int index_bytes,
                        int index_base, void* src_out, void* dst_out, int64_t* perm_out, int64_t* seg_out,
                        int64_t* num_unique, void* stream);
/* gnnb_graph_csr with DEVICE destinations (no synchronisation): the COO -> CSR conversion as an API, for callers that
 * keep working on the device (the reference has no CSR type). */
int gnnb_graph_csr_device(gnnb_graph_t g, int transposed, int32_t* rowptr_dev, int32_t* col_dev, int32_t* eid_dev,
                          void* stream);

/* ------------------------------------------------- neighbour sampling
 * replaces: the edge selection of sample_neighbors(g, nodes, K; dir, replace) (GNNGraphs/src/sampling.jl:68-83), i.e.
 *           adjacency_list(g, nodes; dir, with_eid = true) — a scan of every edge through a Dict on the CPU
 *           (GNNGraphs/src/query.jl:176-198) — plus StatsBase.sample(eidlist[i], k; replace) per node.  The plan's CSR
 *           is that adjacency list, so only the queried rows are touched.
 * nodes: n_nodes DEVICE ids (index_bytes 4|8, index_base 0|1; GNNB_EINDEX if out of range; repeated ids are sampled
 *        independently).  dir = GNNB_DIR_IN samples among the in-edges of each node, GNNB_DIR_OUT among the out-edges.
 * Node j contributes k_j = deg_j (K <= 0), min(K, deg_j) (replace = 0) or K (replace = 1; 0 when deg_j = 0) edges.
 * offsets_dev (n_nodes + 1 int64): running sums of k_j.  eids_dev (NULL = only count): COO positions (index_base-based)
 * of the chosen edges, node after node; capacity = entries available (GNNB_ESIZE if too small).  *total_host = Σ k_j.
 * Draws are counter-based on (seed, j, draw): reproducible per call; without replacement every k-subset of a row is
 * equally likely.  Synchronises the stream. */
int gnnb_sample_neighbors(gnnb_graph_t g, const void* nodes, int64_t n_nodes, int index_bytes, int index_base,
                          int64_t K, int dir, int replace, uint64_t seed, int64_t* offsets_dev, int64_t* eids_dev,
                          int64_t capacity, int64_t* total_host, void* stream);
/* The sampler of ONE query, run on the HOST (no GPU needed): the very function the device kernel calls (compiled for
 * both sides), so that its validity and uniformity can be tested anywhere.  Writes k = k_j positions in [0, deg) to
 * out (capacity entries available) for the counter (seed, j); *k_out = k. */
int gnnb_sample_positions_host(int32_t deg, int64_t K, int replace, uint64_t seed, uint64_t j, int64_t* out,
                               int64_t capacity, int64_t* k_out);

/* ------------------------------------------------------ host-buffer entries
 * The reference-facing call with HOST arrays (what a CPU-array caller of `propagate` has): copies
 * x (and w) to the device, runs the fused pass, copies `out` back; synchronous.  Used for the
 * end-to-end number (bench.py `e2e`).  Same semantics as gnnb_propagate / gnnb_gcn_propagate. */
int gnnb_propagate_host(gnnb_graph_t g, int transposed, int msg, int aggr, const float* x_host,
                        const float* w_host, int64_t D, float* out_host);
int gnnb_gcn_propagate_host(gnnb_graph_t g, int transposed, const float* x_host,
                            const float* w_host, int64_t D, float* out_host);

/* One GCNConv forward (+ backward when dy_host != NULL) on HOST arrays — the call a CPU-array user of the layer makes
 * (bench.py `e2e`): replaces  y = l.σ.(l.weight * (c .* propagate(copy_xj, g', +, xj = x .* c')) .+ l.bias)  and Zygote's
 * pullback (GNNlib/src/layers/conv.jl:14-72 on the `Dout >= Din` branch, default norm_fn, no edge weights).
 * `g` already carries the self loops if the layer adds them.  x_host (Din,N), W_host (Dout,Din) row-major, b_host NULL or
 * Dout, relu 0|1, dy_host (Dout,N) or NULL; outputs y_host (Dout,N), dx_host (Din,N), dW_host (Dout,Din), db_host (Dout or
 * NULL).  Uploads, kernels and downloads run on three streams (x and dy up while y comes down); the device staging lives
 * with the plan.  Pinned host memory makes the copies asynchronous; pageable memory works (staged by the driver).
 * GNNB_EUNSUPPORTED for Dout < Din (use the device entries). */
int gnnb_gcn_conv_step_host(gnnb_graph_t g, const float* x_host, const float* W_host, const float* b_host, int relu,
                            int64_t Din, int64_t Dout, const float* dy_host, float* y_host, float* dx_host,
                            float* dW_host, float* db_host);

/* ------------------------------------------------------------- generators
 * RMAT edge list (ours; the reference has none): Graph500 a,b,c,d = .57,.19,.19,.05,
 * counter-based splitmix64 keyed on (seed, edge id, retry); edges with an endpoint >= N are redrawn;
 * duplicates and self loops kept; generation order.  Writes int64 1-based src/dst DEVICE arrays.
 * The oracle has the bit-identical CPU generator (oracle/gnn_oracle.c: orc_rmat). */
int gnnb_rmat_edges(int64_t num_nodes, int64_t num_edges, uint64_t seed, int64_t* src_dev,
                    int64_t* dst_dev, void* stream);
/* edges [first_edge, first_edge + count) of the same list (the generator is counter-based): a 1 B-edge graph is produced
 * and consumed chunk by chunk (gnnb_shard_builder_add) without ever being resident. */
int gnnb_rmat_edges_range(int64_t num_nodes, int64_t first_edge, int64_t count, uint64_t seed, int64_t* src_dev,
                          int64_t* dst_dev, void* stream);

/* tuning knob for experiments: edges per work chunk of the segmented-reduce kernels (default 128;
 * power of two in [32, 4096]); affects plans created afterwards. */
int gnnb_set_chunk_edges(int chunk);
/* A/B switch for the fused segmented reduce on fp32 rows of 128/256/512 floats (results are bit-identical):
 * 0 = default: the lean work-item kernel (csrc/seglean.cu), taking the plan's per-edge scale stream when there is one;
 * 10 = the lean kernel gathering cs[col] per edge;  12 = seg_reduce_kernel (the round-1 default: register-staged
 * LDG.128, 64-register cap);  5 = the same without the register cap;  1 = TMA-staged: one cp.async.bulk (UBLKCP) per
 * row into a shared-memory ring, mbarrier completion;  13 = the lean pass with rows staged by TMA tile::gather4 (four
 * indexed rows per request into a per-warp shared-memory ring; D = 128 sums).  Measurements: profiles/r1_seg_variants.md, profiles/r2_seg_lean.md. */
int gnnb_set_kernel_variant(int v);

#ifdef __cplusplus
}
#endif
#endif /* GNNB200_H */
