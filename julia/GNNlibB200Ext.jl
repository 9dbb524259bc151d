# GNNlibB200Ext.jl — the Julia side of the drop-in: a GNNlib package extension that routes the message-passing hot
# path to libgnnb200.so (include/gnnb200.h) with `ccall`, in exactly the place where the reference's own CUDA
# extension *disables* its fast path (GNNlib/ext/GNNlibCUDAExt.jl:13-32).
#
# STATUS: source only.  Julia is not installed in the build image or on the GPU box (SURVEY.md §0.3), so this file
# has never been executed; the same entry points are exercised through the Python/ctypes mirror
# (graphneuralnetworks.jl_b200/) whose tests transcribe the reference's.  Registration a maintainer adds to
# GNNlib/Project.toml next to the existing lines (GNNlib/Project.toml:17-23):
#
#     [extensions]
#     GNNlibB200Ext = "CUDA"          # same trigger as GNNlibCUDAExt; methods below are more specific
#
# and `ENV["GNNB200_LIB"]` (or a JLL) pointing at libgnnb200.so.
#
# Method signatures are the reference's own, byte for byte:
#   GNNlib.propagate(::typeof(copy_xj), g::GNNGraph{<:Union{COO_T,SPARSE_T}}, ::typeof(+), xi, xj::AnyCuMatrix, e)
#   (GNNlib/ext/GNNlibCUDAExt.jl:13-16) and its e_mul_xj / w_mul_xj siblings (:21-32), widened to mean/max/min.
module GNNlibB200Ext

using CUDA
using ChainRulesCore
using Statistics: mean
using GNNlib: GNNlib, propagate, copy_xj, e_mul_xj, w_mul_xj
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

# ---- plan cache: graphs are immutable values, keyed on the identity of the (s, t) arrays --------------------------
mutable struct Plan
    h::Ptr{Cvoid}
    function Plan(h)
        p = new(h)
        finalizer(p -> ccall((:gnnb_graph_destroy, LIB), Cint, (Ptr{Cvoid},), p.h), p)
    end
end
const PLANS = IdDict{Any, Plan}()

function plan(g::GNNGraph{<:COO_T})
    s, t = edge_index(g)
    get!(PLANS, s) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:gnnb_graph_create, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Cint, Ptr{Cvoid}),
                    h, pointer(s), pointer(t), g.num_edges, g.num_nodes, g.num_nodes,
                    sizeof(eltype(s)), 1, 1, stream()))
        Plan(h[])
    end
end

const AGGR = IdDict{Any, Cint}(+ => 0, mean => 1, max => 2, min => 3)
const FusedAggr = Union{typeof(+), typeof(mean), typeof(max), typeof(min)}

# ---- the fused forward / pullback (gnnb_propagate, gnnb_propagate_bwd) ---------------------------------------------
function fused_propagate(g, aggr, xj::CuArray{Float32}, w::Union{Nothing, CuVector{Float32}})
    D = length(xj) ÷ size(xj)[end]
    out = similar(xj)
    check(ccall((:gnnb_propagate, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64,
                 CuPtr{Float32}, Ptr{Cvoid}),
                plan(g).h, 0, w === nothing ? 0 : 1, AGGR[aggr], xj, w === nothing ? CU_NULL : w, CU_NULL, CU_NULL, D,
                out, stream()))
    return out
end

function ChainRulesCore.rrule(::typeof(fused_propagate), g, aggr, xj, w)
    out = fused_propagate(g, aggr, xj, w)
    function fused_propagate_pullback(Δ)
        dout = CuArray{Float32}(unthunk(Δ))
        D = length(xj) ÷ size(xj)[end]
        dx = similar(xj)
        dw = w === nothing ? nothing : similar(w)
        check(ccall((:gnnb_propagate_bwd, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
                     CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
                    plan(g).h, w === nothing ? 0 : 1, AGGR[aggr], dout, xj, w === nothing ? CU_NULL : w, CU_NULL, CU_NULL,
                    out, D, dx, dw === nothing ? CU_NULL : dw, stream()))
        return NoTangent(), NoTangent(), NoTangent(), dx, dw === nothing ? NoTangent() : dw
    end
    return out, fused_propagate_pullback
end

## COPY_XJ — replaces GNNlib/ext/GNNlibCUDAExt.jl:13-16 (and adds mean/max/min, 3-D xj)
function GNNlib.propagate(::typeof(copy_xj), g::GNNGraph{<:COO_T}, aggr::FusedAggr, xi,
                          xj::CuArray{Float32}, e)
    fused_propagate(g, aggr, xj, nothing)
end

## E_MUL_XJ with a vector of edge weights — replaces GNNlibCUDAExt.jl:21-24
function GNNlib.propagate(::typeof(e_mul_xj), g::GNNGraph{<:COO_T}, aggr::FusedAggr, xi,
                          xj::CuArray{Float32}, e::CuVector{Float32})
    fused_propagate(g, aggr, xj, e)
end

## W_MUL_XJ with the graph's own weights — replaces GNNlibCUDAExt.jl:29-32
function GNNlib.propagate(::typeof(w_mul_xj), g::GNNGraph{<:COO_T}, aggr::FusedAggr, xi,
                          xj::CuArray{Float32}, e::Nothing)
    fused_propagate(g, aggr, xj, get_edge_weight(g))
end

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

# gcn_conv / gat_conv need no new methods: gcn_conv (GNNlib/src/layers/conv.jl:14-72) reaches the fused kernel through
# the propagate methods above; a further specialisation may call gnnb_gcn_norm + gnnb_gcn_propagate (both 1/sqrt(d)
# scalings folded into the pass) and gnnb_gat_aggregate(+_bwd) exactly as graphneuralnetworks.jl_b200/layers.py does.

end # module
