# GNNlibB200Ext.jl — the Julia side of the drop-in: a GNNlib package extension that routes the message-passing hot
# path to libgnnb200.so (include/gnnb200.h) with `ccall`, in exactly the place where the reference's own CUDA
# extension *disables* its fast path (GNNlib/ext/GNNlibCUDAExt.jl:13-32).
#
# STATUS: source only.  Julia is not installed in the build image or on the GPU box (SURVEY.md §0.3), so this file
# has never been executed; the same entry points are exercised through the Python/ctypes mirror
# (graphneuralnetworks.jl_b200/) whose tests transcribe the reference's.  Registration a maintainer adds to
# GNNlib/Project.toml next to the existing lines (GNNlib/Project.toml:17-23):
#
#     [weakdeps]      CUDA, ChainRulesCore (already a dependency)
#     [extensions]    GNNlibB200Ext = "CUDA"          # same trigger as GNNlibCUDAExt
#
# and `ENV["GNNB200_LIB"]` (or a JLL) pointing at libgnnb200.so.
#
# Dispatch.  GNNlibCUDAExt's three methods are
#   propagate(::typeof(copy_xj|e_mul_xj|w_mul_xj), g::GNNGraph{<:Union{COO_T,SPARSE_T}}, ::typeof(+), xi, xj::AnyCuMatrix, e)
# Both extensions load on the CUDA trigger, and neither those nor the widened methods below (any fused aggregation, any
# rank of xj, Float32 only) are more specific than the other, so the headline call
# `propagate(copy_xj, coo_g, +, xi, ::CuMatrix{Float32}, e)` would be ambiguous.  The methods marked "intersection"
# below carry exactly the intersection signature (COO graph, `+`, CuMatrix{Float32}) and resolve it; a maintainer who
# prefers may instead delete the three methods of GNNlibCUDAExt (INTEGRATION.md says so too).
module GNNlibB200Ext

using CUDA
using ChainRulesCore
using Statistics: mean
using GNNlib: GNNlib, propagate, copy_xj, e_mul_xj, w_mul_xj, expand_srcdst, check_num_nodes
using GNNGraphs: GNNGraphs, GNNGraph, COO_T, edge_index, get_edge_weight

const LIB = get(ENV, "GNNB200_LIB", "libgnnb200")

# ---- status -> the reference's exception types (include/gnnb200.h gnnb_status) ---------------------------------
@inline function check(st::Cint)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:gnnb_last_error, LIB), Cstring, ()))
    (st == 2 || st == 6) && throw(AssertionError(msg))      # GNNB_ESIZE / GNNB_EINDEX  (GNNGraphs/src/utils.jl:1-28)
    st == 1 && throw(ArgumentError(msg))                    # GNNB_EINVAL               (GNNlib/src/layers/conv.jl:3-10)
    error("libgnnb200 status $st: $msg")
end

stream() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)
cuptr(x::Nothing) = CU_NULL
cuptr(x) = pointer(x)

# ---- plan cache ----------------------------------------------------------------------------------------------------
# A graph is an immutable value `(s, t, num_nodes)`; its plan is found through the identity of BOTH index arrays plus the
# sizes.  The table is weak in `s` (a plan dies with the arrays it was built from: the finalizer frees the device CSR)
# and every entry checks `t` by identity through a WeakRef, so two graphs that share `s` but differ in `t` — a reversed
# or rewired graph, GNNGraph(s, t2) — get their own plans.
mutable struct Plan
    h::Ptr{Cvoid}
    loops::Union{Nothing, Plan}          # plan of add_self_loops(g), derived on first use (no second sort)
    function Plan(h)
        p = new(h, nothing)
        finalizer(p -> ccall((:gnnb_graph_destroy, LIB), Cint, (Ptr{Cvoid},), p.h), p)
    end
end
struct PlanEntry
    t::WeakRef
    num_nodes::Int
    num_edges::Int
    plan::Plan
end
const PLANS = WeakKeyDict{Any, Vector{PlanEntry}}()
const PLANS_LOCK = ReentrantLock()

function plan(g::GNNGraph{<:COO_T})
    s, t = edge_index(g)
    lock(PLANS_LOCK) do
        entries = get!(() -> PlanEntry[], PLANS, s)
        filter!(en -> en.t.value !== nothing, entries)
        for en in entries
            (en.t.value === t && en.num_nodes == g.num_nodes && en.num_edges == g.num_edges) && return en.plan
        end
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:gnnb_graph_create, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Cint, Ptr{Cvoid}),
                    h, pointer(s), pointer(t), g.num_edges, g.num_nodes, g.num_nodes,
                    sizeof(eltype(s)), 1, 1, stream()))
        p = Plan(h[])
        push!(entries, PlanEntry(WeakRef(t), g.num_nodes, g.num_edges, p))
        return p
    end
end

# plan of add_self_loops(g) (GNNGraphs/src/transform.jl:12-28): loops appended after the originals
function loop_plan(p::Plan)
    p.loops === nothing || return p.loops
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gnnb_graph_add_self_loops, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ptr{Cvoid}), p.h, h, stream()))
    p.loops = Plan(h[])
end

const AGGR = IdDict{Any, Cint}(+ => 0, mean => 1, max => 2, min => 3)
const FusedAggr = Union{typeof(+), typeof(mean), typeof(max), typeof(min)}

# ---- the fused forward / pullback (gnnb_propagate, gnnb_propagate_bwd) ---------------------------------------------
function fused_propagate(p::Plan, aggr, xj::CuArray{Float32}, w::Union{Nothing, CuVector{Float32}})
    D = length(xj) ÷ size(xj)[end]
    out = similar(xj)
    check(ccall((:gnnb_propagate, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64,
                 CuPtr{Float32}, Ptr{Cvoid}),
                p.h, 0, w === nothing ? 0 : 1, AGGR[aggr], xj, cuptr(w), CU_NULL, CU_NULL, D, out, stream()))
    return out
end

function ChainRulesCore.rrule(::typeof(fused_propagate), p::Plan, aggr, xj, w)
    out = fused_propagate(p, aggr, xj, w)
    function fused_propagate_pullback(Δ)
        dout = CuArray{Float32}(unthunk(Δ))
        D = length(xj) ÷ size(xj)[end]
        dx = similar(xj)
        dw = w === nothing ? nothing : similar(w)
        check(ccall((:gnnb_propagate_bwd, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
                     CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                    p.h, w === nothing ? 0 : 1, AGGR[aggr], dout, xj, cuptr(w), CU_NULL, CU_NULL, out, D, dx, cuptr(dw),
                    stream()))
        return NoTangent(), NoTangent(), NoTangent(), dx, dw === nothing ? NoTangent() : dw
    end
    return out, fused_propagate_pullback
end

## COPY_XJ — replaces GNNlib/ext/GNNlibCUDAExt.jl:13-16 (and adds mean/max/min, 3-D xj)
GNNlib.propagate(::typeof(copy_xj), g::GNNGraph{<:COO_T}, aggr::FusedAggr, xi, xj::CuArray{Float32}, e) =
    fused_propagate(plan(g), aggr, xj, nothing)
GNNlib.propagate(::typeof(copy_xj), g::GNNGraph{<:COO_T}, ::typeof(+), xi, xj::CuMatrix{Float32}, e) =   # intersection
    fused_propagate(plan(g), +, xj, nothing)

## E_MUL_XJ with a vector of edge weights — replaces GNNlibCUDAExt.jl:21-24
GNNlib.propagate(::typeof(e_mul_xj), g::GNNGraph{<:COO_T}, aggr::FusedAggr, xi, xj::CuArray{Float32},
                 e::CuVector{Float32}) = fused_propagate(plan(g), aggr, xj, e)
GNNlib.propagate(::typeof(e_mul_xj), g::GNNGraph{<:COO_T}, ::typeof(+), xi, xj::CuMatrix{Float32},         # intersection
                 e::CuVector{Float32}) = fused_propagate(plan(g), +, xj, e)

## W_MUL_XJ with the graph's own weights — replaces GNNlibCUDAExt.jl:29-32
GNNlib.propagate(::typeof(w_mul_xj), g::GNNGraph{<:COO_T}, aggr::FusedAggr, xi, xj::CuArray{Float32}, e::Nothing) =
    fused_propagate(plan(g), aggr, xj, get_edge_weight(g))
GNNlib.propagate(::typeof(w_mul_xj), g::GNNGraph{<:COO_T}, ::typeof(+), xi, xj::CuMatrix{Float32}, e::Nothing) =   # intersection
    fused_propagate(plan(g), +, xj, get_edge_weight(g))

## softmax_edge_neighbors — replaces GNNlib/src/utils.jl:84-97 on CuArrays
function GNNlib.softmax_edge_neighbors(g::GNNGraph{<:COO_T}, e::CuArray{Float32})
    @assert size(e)[end] == g.num_edges
    K = length(e) ÷ g.num_edges
    out = similar(e)
    check(ccall((:gnnb_softmax_edge_neighbors, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, Int64, CuPtr{Float32}, Ptr{Cvoid}), plan(g).h, e, K, out, stream()))
    return out
end

function ChainRulesCore.rrule(::typeof(GNNlib.softmax_edge_neighbors), g::GNNGraph{<:COO_T}, e::CuArray{Float32})
    α = GNNlib.softmax_edge_neighbors(g, e)
    function softmax_pullback(Δ)
        dα = CuArray{Float32}(unthunk(Δ))
        de = similar(e)
        check(ccall((:gnnb_softmax_edge_neighbors_bwd, LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}, Ptr{Cvoid}),
                    plan(g).h, α, dα, length(e) ÷ g.num_edges, de, stream()))
        return NoTangent(), NoTangent(), de
    end
    return α, softmax_pullback
end

# ---- dense part of a layer: σ.(W * x .+ b) for σ ∈ {identity, relu} (gnnb_linear, gnnb_linear_bwd) -------------------
isrelu(σ) = nameof(σ) === :relu
fusable_σ(σ) = σ === identity || isrelu(σ)

function fused_linear(W::CuMatrix{Float32}, x::CuMatrix{Float32}, b::Union{Nothing, CuVector{Float32}}, relu::Bool)
    Dout, Din = size(W)
    N = size(x, 2)
    # the C entry takes W as the layer stores it in row-major (Dout, Din) terms = Julia's permuted copy
    Wr = permutedims(W)                                  # (Din, Dout) column-major == (Dout, Din) row-major
    y = similar(x, Dout, N)
    check(ccall((:gnnb_linear, LIB), Cint,
                (CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint, Int64, Int64, Int64, CuPtr{Float32}, Ptr{Cvoid}),
                x, Wr, cuptr(b), relu, N, Din, Dout, y, stream()))
    return y
end

function ChainRulesCore.rrule(::typeof(fused_linear), W, x, b, relu::Bool)
    y = fused_linear(W, x, b, relu)
    function fused_linear_pullback(Δ)
        dy = CuArray{Float32}(unthunk(Δ))
        Dout, Din = size(W)
        N = size(x, 2)
        Wr = permutedims(W)
        dx, dWr = similar(x), similar(Wr)
        db = b === nothing ? nothing : similar(b)
        ws = relu ? similar(dy) : nothing
        check(ccall((:gnnb_linear_bwd, LIB), Cint,
                    (CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Cint, Int64, Int64, Int64,
                     CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                    dy, relu ? pointer(y) : CU_NULL, x, Wr, relu, N, Din, Dout, cuptr(ws), dx, dWr, cuptr(db), stream()))
        return NoTangent(), permutedims(dWr), dx, db === nothing ? NoTangent() : db, NoTangent()
    end
    return y, fused_linear_pullback
end

# ---- closing line of a layer that ends in an aggregation: σ.(x .+ b) (gnnb_bias_act, gnnb_bias_act_bwd) ---------------
# one pass forward, one pass backward (mask product + deterministic bias gradient) instead of four broadcasts
function bias_act(x::CuMatrix{Float32}, b::Union{Nothing, CuVector{Float32}}, relu::Bool)
    D, N = size(x)
    y = similar(x)
    check(ccall((:gnnb_bias_act, LIB), Cint,
                (CuPtr{Float32}, CuPtr{Float32}, Cint, Int64, Int64, CuPtr{Float32}, Ptr{Cvoid}),
                x, cuptr(b), relu, N, D, y, stream()))
    return y
end

function ChainRulesCore.rrule(::typeof(bias_act), x, b, relu::Bool)
    y = bias_act(x, b, relu)
    function bias_act_pullback(Δ)
        dy = CuArray{Float32}(unthunk(Δ))
        D, N = size(dy)
        dpre = relu ? similar(dy) : dy
        db = b === nothing ? nothing : similar(b)
        (relu || db !== nothing) &&
            check(ccall((:gnnb_bias_act_bwd, LIB), Cint,
                        (CuPtr{Float32}, CuPtr{Float32}, Cint, Int64, Int64, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                        dy, relu ? pointer(y) : CU_NULL, relu, N, D, relu ? pointer(dpre) : CU_NULL, cuptr(db), stream()))
        return NoTangent(), dpre, db === nothing ? NoTangent() : db, NoTangent()
    end
    return y, bias_act_pullback
end

closing(l, y::CuMatrix{Float32}) =
    (fusable_σ(l.σ) && size(y, 1) % 4 == 0 && size(y, 1) <= 1024 && (l.bias === false || l.bias isa CuVector{Float32})) ?
    bias_act(y, l.bias === false ? nothing : l.bias, isrelu(l.σ)) : l.σ.(y .+ l.bias)

dense(l, W, x, with_bias_act::Bool) = begin
    b = (with_bias_act && l.bias isa CuVector{Float32}) ? l.bias : nothing
    σ = with_bias_act ? l.σ : identity
    if fusable_σ(σ) && size(W, 1) % 4 == 0 && size(W, 2) % 4 == 0 && (l.bias === false || l.bias isa CuVector{Float32} || !with_bias_act)
        fused_linear(W, x, b, isrelu(σ))
    else
        with_bias_act ? σ.(W * x .+ l.bias) : W * x
    end
end

# ---- gcn_conv fast path (GNNlib/src/layers/conv.jl:14-72; callers GraphNeuralNetworks/src/layers/conv.jl:103, GNNLux :137)
# c .* propagate(copy_xj | e_mul_xj, g', +, xj = x .* c') in ONE pass over the self-loop plan, both scalings folded into the
# gather and the store (gnnb_gcn_propagate), and its pullback (the same kernel on the CSR-by-source plan).
function gcn_core(p::Plan, x::CuMatrix{Float32}, c::CuVector{Float32})
    out = similar(x)
    check(ccall((:gnnb_gcn_propagate, LIB), Cint,
                (Ptr{Cvoid}, Cint, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}, Ptr{Cvoid}),
                p.h, 0, x, CU_NULL, c, size(x, 1), out, stream()))
    return out
end
function ChainRulesCore.rrule(::typeof(gcn_core), p::Plan, x, c)
    out = gcn_core(p, x, c)
    function gcn_core_pullback(Δ)
        dout = CuArray{Float32}(unthunk(Δ))
        dx = similar(x)
        check(ccall((:gnnb_gcn_propagate, LIB), Cint,
                    (Ptr{Cvoid}, Cint, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}, Ptr{Cvoid}),
                    p.h, 1, dout, CU_NULL, c, size(x, 1), dx, stream()))
        return NoTangent(), NoTangent(), dx, NoTangent()       # c is a function of the graph only (unweighted)
    end
    return out, gcn_core_pullback
end

function in_degree(p::Plan, n::Integer)
    d = CUDA.zeros(Float32, n)
    check(ccall((:gnnb_degree, LIB), Cint, (Ptr{Cvoid}, Cint, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                p.h, 1, CU_NULL, d, stream()))
    return d
end
ChainRulesCore.@non_differentiable in_degree(::Any...)
ChainRulesCore.@non_differentiable plan(::Any...)
ChainRulesCore.@non_differentiable loop_plan(::Any...)

function GNNlib.gcn_conv(l, g::GNNGraph{<:COO_T}, x::CuMatrix{Float32}, edge_weight::Nothing, norm_fn::F,
                         conv_weight::Union{Nothing, CuMatrix{Float32}}) where {F}
    (l.use_edge_weight && get_edge_weight(g) !== nothing) &&            # weighted graphs: the reference's own body, whose
        return invoke(GNNlib.gcn_conv, Tuple{Any, GNNlib.AbstractGNNGraph, Any, Nothing, F, typeof(conv_weight)},
                      l, g, x, edge_weight, norm_fn, conv_weight)      # propagate calls reach the fused methods above
    weight = conv_weight === nothing ? l.weight : conv_weight
    size(weight) == size(l.weight) ||
        throw(ArgumentError("The weight matrix has the wrong size. Expected $(size(l.weight)) but got $(size(weight))"))
    check_num_nodes(g, x)
    Dout, Din = size(weight)
    p = l.add_self_loops ? loop_plan(plan(g)) : plan(g)
    Dout < Din && (x = dense(l, weight, x, false))                       # multiply before convolution (conv.jl:36-40)
    c = norm_fn(in_degree(p, g.num_nodes))                               # conv.jl:52-57, any norm_fn
    x = gcn_core(p, x, c)
    Dout >= Din && return dense(l, weight, x, true)                      # σ.(W * x .+ b), bias/relu in the GEMM epilogue
    return l.σ.(x .+ l.bias)
end

# ---- gat_conv fast path (GNNlib/src/layers/conv.jl:112-167; callers GraphNeuralNetworks/src/layers/conv.jl:346) ------
# logits -> leakyrelu -> neighbourhood softmax -> α-weighted sum of Wx rows in one pass (gnnb_gat_aggregate), no (C,H,E)
# tensors; the pullback recomputes α from the per-target statistics (gnnb_gat_aggregate_bwd).
function gat_aggregate(p::Plan, Wx::CuArray{Float32, 3}, el::CuMatrix{Float32}, er::CuMatrix{Float32}, slope::Float32)
    C, H, N = size(Wx)
    out = similar(Wx)
    check(ccall((:gnnb_gat_aggregate, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int64, Cfloat, CuPtr{Float32},
                 CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                p.h, Wx, el, er, C, H, slope, out, CU_NULL, CU_NULL, CU_NULL, stream()))
    return out
end
function ChainRulesCore.rrule(::typeof(gat_aggregate), p::Plan, Wx, el, er, slope)
    C, H, N = size(Wx)
    out = similar(Wx)
    smax, ssum = similar(el), similar(el)
    check(ccall((:gnnb_gat_aggregate, LIB), Cint,
                (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int64, Cfloat, CuPtr{Float32},
                 CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                p.h, Wx, el, er, C, H, slope, out, CU_NULL, smax, ssum, stream()))
    function gat_aggregate_pullback(Δ)
        dout = CuArray{Float32}(unthunk(Δ))
        dWx, del, der = similar(Wx), similar(el), similar(er)
        check(ccall((:gnnb_gat_aggregate_bwd, LIB), Cint,
                    (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
                     CuPtr{Float32}, CuPtr{Float32}, Int64, Int64, Cfloat, CuPtr{Float32}, CuPtr{Float32},
                     CuPtr{Float32}, Ptr{Cvoid}),
                    p.h, Wx, el, er, smax, ssum, out, dout, C, H, slope, dWx, del, der, stream()))
        return NoTangent(), NoTangent(), dWx, del, der, NoTangent()
    end
    return out, gat_aggregate_pullback
end

function GNNlib.gat_conv(l, g::GNNGraph{<:COO_T}, x::CuMatrix{Float32}, e::Nothing = nothing)
    (l.dense_e === nothing && iszero(l.dropout)) ||                      # edge features / dropout: the reference's body
        return invoke(GNNlib.gat_conv, Tuple{Any, GNNlib.AbstractGNNGraph, Any, Nothing}, l, g, x, e)
    check_num_nodes(g, x)
    p = l.add_self_loops ? loop_plan(plan(g)) : plan(g)
    _, chout = l.channel
    heads = l.heads
    Wx = reshape(l.dense_x(x), chout, heads, :)                          # conv.jl:128-129 (Dense without bias, :336)
    ai, aj = l.a[1:chout, :], l.a[(chout + 1):(2chout), :]               # rows 1..C pair with the target, C+1..2C with the source (:157)
    el = dropdims(sum(reshape(ai, chout, heads, 1) .* Wx, dims = 1), dims = 1)    # (H, N)
    er = dropdims(sum(reshape(aj, chout, heads, 1) .* Wx, dims = 1), dims = 1)
    y = gat_aggregate(p, Wx, el, er, Float32(l.negative_slope))
    l.concat || (y = mean(y, dims = 2))
    y = reshape(y, :, size(y, 3))
    return closing(l, y)                                                 # conv.jl:149, one pass
end

# ---- sage_conv fast path (GNNlib/src/layers/conv.jl:277-283; caller GraphNeuralNetworks/src/layers/conv.jl:787) ------
# σ.(W * [x_i ; m_i] .+ b) without the (2D, N) vcat temporary: the two column blocks of W hit x and m separately.
function GNNlib.sage_conv(l, g::GNNGraph{<:COO_T}, x::CuMatrix{Float32})
    check_num_nodes(g, x)
    xj, xi = expand_srcdst(g, x)
    m = propagate(copy_xj, g, l.aggr, xj = xj)                           # the fused mean / sum / max kernel
    Din = size(xi, 1)
    W1, W2 = l.weight[:, 1:Din], l.weight[:, (Din + 1):end]
    return l.σ.(W1 * xi .+ W2 * m .+ l.bias)
end

## sort_edge_index on device arrays — replaces GNNGraphs/ext/GNNGraphsCUDAExt.jl:24-30 (copy to the CPU, sort, copy back)
function GNNGraphs.sort_edge_index(u::CuVector{T}, v::CuVector{T}) where {T <: Union{Int32, Int64}}
    @assert length(u) == length(v)
    uo, vo = similar(u), similar(v)
    isempty(u) && return uo, vo
    hi = max(maximum(u), maximum(v))
    check(ccall((:gnnb_sort_edge_index, LIB), Cint,
                (CuPtr{T}, CuPtr{T}, Int64, Int64, Cint, CuPtr{T}, CuPtr{T}, CuPtr{Int64}, Ptr{Cvoid}),
                u, v, length(u), hi, sizeof(T), uo, vo, CU_NULL, stream()))
    return uo, vo
end

## remove_multi_edges on a device COO graph — the index half of GNNGraphs/src/transform.jl:157-190 in one call; the
## `_scatter(aggr, ·, idxs)` of the weights / edge features stays the reference's own line (NNlib.scatter on CuArrays).
function coalesce_edge_index(s::CuVector{T}, t::CuVector{T}, n::Integer) where {T <: Union{Int32, Int64}}
    E = length(s)
    so, to = similar(s), similar(t)
    perm, idxs = CUDA.zeros(Int64, E), CUDA.zeros(Int64, E)
    nu = Ref{Int64}(0)
    check(ccall((:gnnb_coalesce_edges, LIB), Cint,
                (CuPtr{T}, CuPtr{T}, Int64, Int64, Cint, Cint, CuPtr{T}, CuPtr{T}, CuPtr{Int64}, CuPtr{Int64},
                 Ref{Int64}, Ptr{Cvoid}),
                s, t, E, n, sizeof(T), 1, so, to, perm, idxs, nu, stream()))
    return so[1:nu[]], to[1:nu[]], perm .+ 1, idxs          # 1-based permutation, segment id of every sorted edge
end

end # module
