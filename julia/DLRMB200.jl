# DLRMB200.jl -- Julia host glue for libdlrm_b200.so (the C ABI in include/dlrm_b200.h).
#
# STATUS: written against the reference's API, NOT EXECUTED.  There is no `julia` binary in the
# build image or on the GPU box, and DLRM.jl's path dependencies (EmbeddingTables, OneDNN,
# CachedArrays; Manifest.toml:79-83,246-250,522-526) are not vendored.  Every ccall below is
# exercised instead through the Python ctypes binding (dlrm_jl_b200/_lib.py), which binds the same
# symbols with the same argument lists; tests/test_abi_cpu.py checks that binding against the header.
#
# What it replaces in DLRM.jl (paths relative to the DLRM.jl checkout):
#   * embedding_constructor / embedding_allocator hooks of `dlrm`  (src/model/model.jl:185-206)
#       -> B200Tables(data::Vector{Matrix{Float32}})          HBM-resident tables
#   * maplookup(strategy, tables, sparse) + its pullback          (src/model/model.jl:161)
#   * (dot::DotInteraction)(x, ys) + rrule                        (src/model/interact.jl:394-447)
#   * EmbeddingTables.update!(opt, tables, grads, indexers; ...)  (src/train/train.jl:283-290)
#   * Array(table) read-back used by validate_embeddings          (src/validation.jl:138)
# The model in DLRM.jl is CPU-resident (OneDNN MLPs), so this glue uses the `_host` entry points:
# Julia arrays go in and come out, the copies happen inside the call.  Column-major Julia arrays
# are passed as they are: a `D x N` Julia matrix is the C array [N][D] the library expects.
module DLRMB200

import ChainRulesCore
import ChainRulesCore: NoTangent
import Flux

const libdlrm_b200 = get(ENV, "DLRM_B200_LIB",
                         joinpath(@__DIR__, "..", "dlrm_jl_b200", "lib", "libdlrm_b200.so"))

struct B200Error <: Exception
    status::Int32
    msg::String
end
Base.showerror(io::IO, e::B200Error) = print(io, "libdlrm_b200 status ", e.status, ": ", e.msg)

function check(rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dlrmb_last_error, libdlrm_b200), Cstring, ()))
    throw(B200Error(rc, msg))
end

#####
##### Tables: Vector{SimpleEmbedding{Static{D}}}  ->  one opaque device handle
#####

mutable struct B200Tables
    handle::Ptr{Cvoid}
    rows::Vector{Int64}
    featuresize::Int
    max_lookups::Int
end

"""
    B200Tables(data; max_lookups, device = 0)

`data[k]` is the `D x nrows_k` matrix the reference hands to `SimpleEmbedding{Static{D}}(data)`
(src/data/criteo.jl:490).  Use as `embedding_constructor` by building all tables first and
constructing once; `max_lookups` is the largest `batchsize * lookups_per_sample` of any batch.
"""
function B200Tables(data::Vector{<:AbstractMatrix{Float32}}; max_lookups::Integer, device::Integer = 0)
    D = size(first(data), 1)
    rows = Int64[size(m, 2) for m in data]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dlrmb_tables_create, libdlrm_b200), Int32,
                (Int32, Int32, Ptr{Int64}, Int32, Int64, Ptr{Ptr{Cvoid}}),
                device, length(rows), rows, D, max_lookups, out))
    t = B200Tables(out[], rows, D, max_lookups)
    finalizer(t) do x
        x.handle == C_NULL || ccall((:dlrmb_tables_destroy, libdlrm_b200), Int32, (Ptr{Cvoid},), x.handle)
        x.handle = C_NULL
    end
    for (k, m) in enumerate(data)
        A = Matrix{Float32}(m)   # contiguous D x nrows == C [nrows][D]
        check(ccall((:dlrmb_tables_upload, libdlrm_b200), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}),
                    t.handle, k - 1, A))
    end
    return t
end

Base.length(t::B200Tables) = length(t.rows)

"`Array(tables, k)`: the `D x nrows_k` matrix of table k (validate_embeddings, src/validation.jl:138)."
function Base.Array(t::B200Tables, k::Integer)
    A = Matrix{Float32}(undef, t.featuresize, t.rows[k])
    check(ccall((:dlrmb_tables_download, libdlrm_b200), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}),
                t.handle, k - 1, A))
    return A
end

#####
##### Index containers -> table-major [ntab][B][P]
#####

# `sparse` is DACLoader's B x ntab UInt32 matrix (src/data/criteo.jl:320-326) or a vector of
# per-table index vectors / P x B matrices (src/data/criteo.jl:551-557).  Both are already the
# table-major layout the library wants once concatenated; indices stay 1-based (idx_base = 1).
_pack(sparse::AbstractMatrix{<:Integer}) = (Matrix(sparse), 1)                        # P = 1
function _pack(sparse::AbstractVector)
    P = ndims(first(sparse)) == 1 ? 1 : size(first(sparse), 1)
    return (reduce(hcat, [vec(s) for s in sparse]), P)                                # (P*B) x ntab
end
_idxbytes(::AbstractArray{T}) where {T} = Int32(sizeof(T))

#####
##### maplookup + pullback
#####

struct PreallocationStrategy
    prependrows::Int
end
PreallocationStrategy() = PreallocationStrategy(0)

"Sparse gradient of one lookup call: the whole pooled-gradient matrix plus the packed indices."
struct B200SparseUpdate{I<:AbstractMatrix}
    delta::Matrix{Float32}      # (prependrows + ntab*D) x B == C [B][slots][D]
    indices::I
    P::Int
    slot0::Int
end

function maplookup(strategy::PreallocationStrategy, tables::B200Tables, sparse)
    idx, P = _pack(sparse)
    B = div(size(idx, 1), P)
    D = tables.featuresize
    @assert iszero(mod(strategy.prependrows, D))
    slot0 = div(strategy.prependrows, D)
    slots = slot0 + length(tables)
    out = zeros(Float32, slots * D, B)
    check(ccall((:dlrmb_embedding_fwd_host, libdlrm_b200), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float32}, Int32, Int32),
                tables.handle, idx, _idxbytes(idx), 1, B, P, out, slots, slot0))
    return out
end

function ChainRulesCore.rrule(::typeof(maplookup), strategy::PreallocationStrategy, tables::B200Tables, sparse)
    out = maplookup(strategy, tables, sparse)
    idx, P = _pack(sparse)
    slot0 = div(strategy.prependrows, tables.featuresize)
    function maplookup_pullback(Δ)
        # same tuple shape as the reference: (nothing, nothing, updates, nothing)
        # (test/model/embedding_update.jl:36-40)
        return (NoTangent(), NoTangent(), B200SparseUpdate(Matrix{Float32}(Δ), idx, P, slot0), NoTangent())
    end
    return out, maplookup_pullback
end

#####
##### DotInteraction
#####

struct B200DotInteraction
    tables::B200Tables          # owns the staging buffers / stream the host entry points use
    pad_to_mul::Int             # POST_INTERACTION_PAD_TO_MUL, src/model/model.jl:32
end
B200DotInteraction(tables::B200Tables) = B200DotInteraction(tables, 1)

_width(F, d, m) = m * cld(d + div(F * (F - 1), 2), m)

function (dot::B200DotInteraction)(x::AbstractMatrix{Float32}, ys::AbstractMatrix{Float32})
    d, B = size(x)
    F = div(size(ys, 1), d)
    out = Matrix{Float32}(undef, _width(F, d, dot.pad_to_mul), B)
    # x is copied into slot 0 of ys inside the kernel (fast_vcat, src/model/interact.jl:271-281)
    check(ccall((:dlrmb_interaction_fwd_host, libdlrm_b200), Int32,
                (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32, Int32, Int32, Int32, Ptr{Float32}),
                dot.tables.handle, ys, Matrix{Float32}(x), B, F, d, dot.pad_to_mul, out))
    return out
end

function ChainRulesCore.rrule(dot::B200DotInteraction, x::AbstractMatrix{Float32}, ys::AbstractMatrix{Float32})
    out = dot(x, ys)                 # ys now holds x in slot 0: it is the saved `t`
    d, B = size(x)
    F = div(size(ys, 1), d)
    function dot_pullback(Δ)
        dT = similar(ys)
        dx = Matrix{Float32}(undef, d, B)
        check(ccall((:dlrmb_interaction_bwd_host, libdlrm_b200), Int32,
                    (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32, Int32, Int32, Int32, Ptr{Float32}, Ptr{Float32}),
                    dot.tables.handle, Matrix{Float32}(Δ), ys, B, F, d, dot.pad_to_mul, dT, dx))
        return (NoTangent(), dx, dT)   # dy is the whole (d*F) x B matrix (src/model/interact.jl:428-435)
    end
    return out, dot_pullback
end

#####
##### update!
#####

"`EmbeddingTables.update!(opt, tables, grads, indexers; num_splits, nthreads)` for B200 tables."
function update!(opt::Flux.Descent, tables::B200Tables, grads::B200SparseUpdate, indexers = nothing;
                 num_splits = 8, nthreads = 12)
    slots, B = div(size(grads.delta, 1), tables.featuresize), size(grads.delta, 2)
    check(ccall((:dlrmb_embedding_bwd_sgd_host, libdlrm_b200), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float32}, Int32, Int32, Float32),
                tables.handle, grads.indices, _idxbytes(grads.indices), 1, B, grads.P, grads.delta,
                slots, grads.slot0, Float32(opt.eta)))
    return nothing
end

end # module
