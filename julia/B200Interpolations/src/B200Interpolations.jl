"""
    B200Interpolations

Host-side glue that makes `libb200itp.so` (hand-written sm_100a kernels behind the C ABI of `include/b200itp.h`) a
drop-in for the B-spline hot path of Interpolations.jl v0.16: `interpolate(A, BSpline(...))` on a device array,
`scale`, `extrapolate`, and the broadcasts `itp.(xs...)`, `Interpolations.gradient.(Ref(itp), xs...)`,
`Interpolations.hessian.(Ref(itp), xs...)`.  No reference source changes: both seams are existing dispatch points —

  * prefilter:  `AxisAlgorithms.A_ldiv_B_md!(dest, M, src, dim, b)` (src/filter1d.jl:8-85, called from
                src/b-splines/prefiltering.jl:49-59,108-122) and `copy_with_padding` (prefiltering.jl:17-28),
                dispatched on the array type `B200Array`;
  * evaluation: `Base.BroadcastStyle(root_storage_type(typeof(itp)))` (src/gpu_support.jl:52-80) selects `B200Style`
                for interpolants whose coefficients live in a `B200Array`, and `Base.copy(::Broadcasted{<:B200Style})`
                below unpacks the wrapper chain into a `b200itp_desc` and `ccall`s the library.

Every `ccall` signature in this file is checked against `include/b200itp.h` by `tests/test_julia_glue.py` (arity and
argument kinds), and the field order of `Desc` / `AxisSystem` against the C structs.  Julia is not part of the build
image, so this file is reviewed, not executed, there; `test/runtests.jl` covers the cases of the reference's own
array-type-generic testset (`test/gpu_support.jl`) with `==`.
"""
module B200Interpolations

using Adapt
using LinearAlgebra
using OffsetArrays
using StaticArrays
using Interpolations
using Interpolations: AbstractInterpolation, BSplineInterpolation, ScaledInterpolation, Extrapolation,
    FilledExtrapolation, BSpline, Degree, Constant, Linear, Quadratic, Cubic, NoInterp, Nearest, Previous, Next,
    BoundaryCondition, Throw, Flat, Line, Free, Periodic, Reflect, InPlace, InPlaceQ, OnGrid, OnCell,
    lbounds, ubounds, itpflag, padded_axes
import Interpolations: copy_with_padding
import AxisAlgorithms: A_ldiv_B_md!
using WoodburyMatrices: Woodbury

export B200Array, b200, device_synchronize, Comm, replicate, interpolate_sharded, evaluate_sharded

# ------------------------------------------------------------------------------------------------ library
const libb200itp = get(ENV, "B200ITP_LIB", "libb200itp")

# codes of include/b200itp.h
const F32, F64 = Cint(0), Cint(1)
const TAG_EXACT, TAG_F32, TAG_F64 = Cint(0), Cint(1), Cint(2)
const ETP_NONE, ETP_FLAGS, ETP_FILL = Cint(0), Cint(1), Cint(2)
const MAX_DIMS = 4

function last_error()
    buf = Vector{UInt8}(undef, 1024)
    n = ccall((:b200itp_last_error, libb200itp), Csize_t, (Ptr{UInt8}, Csize_t), buf, 1024)
    return String(buf[1:min(Int(n), 1023)])
end

"Error convention of the C ABI (INTEGRATION.md §5): 0 ok, > 0 unsupported / bad argument, < 0 CUDA failure."
function check(rc::Integer)
    rc == 0 && return nothing
    msg = last_error()
    rc == 4 && throw(DimensionMismatch(msg))          # B200ITP_E_SIZE: the messages of filter1d.jl:9,11
    rc > 0 && throw(ArgumentError(msg))
    error("libb200itp: CUDA failure $(rc): $(msg)")
end

dtype_code(::Type{Float32}) = F32
dtype_code(::Type{Float64}) = F64
dtype_code(::Type{T}) where {T} = throw(ArgumentError("libb200itp handles Float32 and Float64 arrays, got $T"))

# ------------------------------------------------------------------------------------------------ device array
"""
    B200Array{T,N}

Dense column-major array in the HBM of one B200 (`dev`), owned by Julia (finalizer frees it).  Deliberately minimal:
it is the storage type the interpolation objects are adapted to, not a general GPU array library.  Scalar indexing is
not provided (like `JLArrays.allowscalar(false)`).
"""
mutable struct B200Array{T,N} <: AbstractArray{T,N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N,Int}
    dev::Cint
    function B200Array{T,N}(::UndefInitializer, dims::NTuple{N,Int}; dev::Integer=0) where {T,N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        nbytes = max(prod(dims) * sizeof(T), 1)
        check(ccall((:b200itp_malloc, libb200itp), Cint, (Cint, Csize_t, Ptr{Ptr{Cvoid}}), dev, nbytes, p))
        a = new{T,N}(p[], dims, Cint(dev))
        finalizer(a) do x
            ccall((:b200itp_free, libb200itp), Cint, (Cint, Ptr{Cvoid}), x.dev, x.ptr)
        end
        return a
    end
end
B200Array{T}(u::UndefInitializer, dims::NTuple{N,Int}; dev::Integer=0) where {T,N} = B200Array{T,N}(u, dims; dev=dev)
B200Array{T}(u::UndefInitializer, dims::Int...; dev::Integer=0) where {T} = B200Array{T}(u, dims; dev=dev)

Base.size(a::B200Array) = a.dims
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
Base.getindex(::B200Array, ::Int...) = error("scalar indexing of a B200Array is disallowed; use Array(a)")
Base.setindex!(::B200Array, _, ::Int...) = error("scalar indexing of a B200Array is disallowed")
Base.similar(a::B200Array, ::Type{T}, dims::Dims{N}) where {T,N} = B200Array{T,N}(undef, dims; dev=a.dev)
Base.unsafe_convert(::Type{Ptr{Cvoid}}, a::B200Array) = a.ptr
Base.show(io::IO, a::B200Array{T,N}) where {T,N} = print(io, "B200Array{$T,$N}", size(a), " on device ", a.dev)
Base.show(io::IO, ::MIME"text/plain", a::B200Array) = show(io, a)

function device_synchronize(dev::Integer=0)
    check(ccall((:b200itp_sync, libb200itp), Cint, (Cint, Ptr{Cvoid}), dev, C_NULL))
end

"Upload: `B200Array(A)`"
function B200Array(A::Array{T,N}; dev::Integer=0) where {T,N}
    d = B200Array{T,N}(undef, size(A); dev=dev)
    GC.@preserve A check(ccall((:b200itp_memcpy_h2d, libb200itp), Cint,
                               (Cint, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                               d.dev, d.ptr, pointer(A), sizeof(A), C_NULL))
    return d
end
B200Array(A::AbstractArray; dev::Integer=0) = B200Array(collect(A); dev=dev)

"Download: `Array(a)` / `collect(a)`"
function Base.Array(a::B200Array{T,N}) where {T,N}
    h = Array{T,N}(undef, size(a))
    GC.@preserve h check(ccall((:b200itp_memcpy_d2h, libb200itp), Cint,
                               (Cint, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                               a.dev, pointer(h), a.ptr, sizeof(h), C_NULL))
    device_synchronize(a.dev)
    return h
end
Base.collect(a::B200Array) = Array(a)
# `first(A)` is the one scalar read the reference makes on a coefficient array: `BSplineInterpolation(TWeights, A, it, axs)`
# fixes the interpolant's element type with `typeof(zero(TWeights) * first(A))` (b-splines.jl:104-108)
function Base.first(a::B200Array{T}) where {T}
    isempty(a) && throw(BoundsError(a, 1))
    h = Ref{T}()
    check(ccall((:b200itp_memcpy_d2h, libb200itp), Cint, (Cint, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                a.dev, h, a.ptr, sizeof(T), C_NULL))
    device_synchronize(a.dev)
    return h[]
end
Base.first(a::OffsetArray{T,N,<:B200Array{T,N}}) where {T,N} = first(parent(a))
Base.copy(a::B200Array) = B200Array(Array(a); dev=a.dev)   # (device-to-device copies go through the host: not on the hot path)
function Base.copyto!(dest::B200Array{T,N}, src::B200Array{T,N}) where {T,N}
    size(dest) == size(src) || throw(DimensionMismatch("copyto!: sizes differ"))
    dest.ptr == src.ptr && return dest
    h = Array(src)
    GC.@preserve h check(ccall((:b200itp_memcpy_h2d, libb200itp), Cint,
                               (Cint, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                               dest.dev, dest.ptr, pointer(h), sizeof(h), C_NULL))
    return dest
end

# Adapt: `adapt(B200Array, itp)` rebuilds the wrapper chain around the adapted coefficients with the reference's own
# `adapt_structure` methods (src/gpu_support.jl:4-8,32-48); only the storage rule is ours.
Adapt.adapt_storage(::Type{<:B200Array}, A::Array{T}) where {T<:Union{Float32,Float64}} = B200Array(A)
Adapt.adapt_storage(::Type{<:B200Array}, A::B200Array) = A
"`b200(x)`: move an array or an interpolation object to the B200 (the counterpart of `JLArrays.jl`'s `jl`)."
b200(x) = adapt(B200Array, x)

const DeviceCoefs{T,N} = Union{B200Array{T,N},OffsetArray{T,N,<:B200Array{T,N}}}
devarray(a::B200Array) = a
devarray(a::OffsetArray) = parent(a)

# ------------------------------------------------------------------------------------------------ prefilter seam
"`b200itp_axis_system` (include/b200itp.h): the factorisation `prefiltering_system` returned, host pointers."
struct AxisSystem
    dtype::Int32
    k::Int32
    n::Int64
    dl::Ptr{Cvoid}
    d::Ptr{Cvoid}
    du::Ptr{Cvoid}
    urow::Ptr{Int32}
    vcol::Ptr{Int32}
    Cp::Ptr{Cvoid}
end

# `M` is what `prefiltering_system` returns (cubic.jl:101-264, quadratic.jl:71-189): `lut!(dl, d, du)`, a
# `LU{T,Tridiagonal}` whose `factors` hold the multipliers / pivots / super-diagonal (b-splines.jl:257), bare or inside
# `Woodbury(A, U, C, V)`; U has unit columns, V unit rows (`sparse_factors`), `Cp` is the capacitance matrix.
lu_factors(F::LU) = F.factors
unit_index(v) = Int32(findfirst(!iszero, v))

function A_ldiv_B_md!(dest::B200Array{T,N}, M::Union{LU,Woodbury}, src::B200Array{T,N}, dim::Integer,
                      b::AbstractVector) where {T<:Union{Float32,Float64},N}
    1 <= dim <= max(N, 1) || throw(DimensionMismatch("The chosen dimension $dim is larger than $N"))   # filter1d.jl:9
    n = size(M, 1)
    n == size(src, dim) || throw(DimensionMismatch("Sizes $n and $(size(src, dim)) do not match"))     # filter1d.jl:11
    size(dest) == size(src) || throw(DimensionMismatch("Sizes $(size(dest)) and $(size(src)) do not match"))
    all(iszero, b) || throw(ArgumentError("B200Interpolations: only the zero offset of prefilter! is supported"))
    copyto!(dest, src)
    F = M isa Woodbury ? M.A : M
    f = lu_factors(F)
    dl, d, du = Vector{T}(f.dl), Vector{T}(f.d), Vector{T}(f.du)
    k = M isa Woodbury ? size(M.U, 2) : 0
    urow = Int32[unit_index(view(M.U, :, j)) for j in 1:k]
    vcol = Int32[unit_index(view(M.V, j, :)) for j in 1:k]
    Cp = k == 0 ? T[] : vec(Matrix{T}(M.Cp))
    dims = Int64[size(dest)...]
    GC.@preserve dl d du urow vcol Cp dims begin
        sys = AxisSystem(dtype_code(T), Int32(k), Int64(n), pointer(dl), pointer(d), pointer(du),
                         k == 0 ? Ptr{Int32}(C_NULL) : pointer(urow), k == 0 ? Ptr{Int32}(C_NULL) : pointer(vcol),
                         k == 0 ? C_NULL : pointer(Cp))
        check(ccall((:b200itp_prefilter_axis, libb200itp), Cint,
                    (Cint, Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Cint, Ref{AxisSystem}, Ptr{Cvoid}),
                    dest.dev, dest.ptr, dtype_code(T), N, dims, dim - 1, Ref(sys), C_NULL))
    end
    return dest
end
A_ldiv_B_md!(dest::B200Array, ::Nothing, src::B200Array, dim::Integer, b::AbstractVector) = copyto!(dest, src)  # filter1d.jl:6

# `copy_with_padding(TC, A, it)` (prefiltering.jl:17-28): zero border of one cell on every padded axis, interior = A.
# (`it` carries the reference method's own constraint: with it untyped the two methods would be ambiguous)
function copy_with_padding(::Type{TC}, A::B200Array{T,N},
                           it::Interpolations.DimSpec{Interpolations.InterpolationType}) where {TC,T,N}
    TC === T || throw(ArgumentError("B200Interpolations: coefficient type $TC differs from the array type $T"))
    indsA = axes(A)
    indspad = padded_axes(indsA, it)
    pad = Int32[first(a) - first(p) for (a, p) in zip(indsA, indspad)]
    coefs = B200Array{T,N}(undef, map(length, indspad); dev=A.dev)
    ssz = Int64[size(A)...]
    GC.@preserve ssz pad check(ccall((:b200itp_copy_with_padding, libb200itp), Cint,
                                     (Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int32}, Ptr{Cvoid}),
                                     A.dev, A.ptr, coefs.ptr, dtype_code(T), N, ssz, pad, C_NULL))
    return indspad == indsA ? coefs : OffsetArray(coefs, indspad)
end

# ------------------------------------------------------------------------------------------------ descriptor
"`b200itp_desc` (include/b200itp.h), field for field; `NTuple`s are the inline C arrays."
struct Desc
    ndims::Int32
    coef_dtype::Int32
    coord_dtype::Int32
    etp_pair_mask::Int32
    size::NTuple{4,Int64}
    parent_len::NTuple{4,Int64}
    coef_first::NTuple{4,Int32}
    degree::NTuple{4,Int32}
    bc::NTuple{4,Int32}
    gridtype::NTuple{4,Int32}
    inner_lb::NTuple{4,Float64}
    inner_ub::NTuple{4,Float64}
    inner_tag::NTuple{4,Int32}
    has_scale::Int32
    range_tag::NTuple{4,Int32}
    range_first::NTuple{4,Float64}
    range_step::NTuple{4,Float64}
    outer_lb::NTuple{4,Float64}
    outer_ub::NTuple{4,Float64}
    outer_tag::NTuple{4,Int32}
    etp_kind::Int32
    etp_lo::NTuple{4,Int32}
    etp_hi::NTuple{4,Int32}
    fill_tag::Int32
    fill_value::Float64
    parent_first::NTuple{4,Int32}
end

pad4(t, z) = ntuple(i -> i <= length(t) ? t[i] : z, 4)

# integer codes (include/b200itp.h)
degree_code(::Linear) = Int32(1)
degree_code(::Quadratic) = Int32(2)
degree_code(::Cubic) = Int32(3)
degree_code(::Constant{Nearest}) = Int32(4)
degree_code(::Constant{Previous}) = Int32(5)
degree_code(::Constant{Next}) = Int32(6)
bc_code(::Throw) = Int32(0)
bc_code(::Flat) = Int32(1)
bc_code(::Line) = Int32(2)
bc_code(::Free) = Int32(3)
bc_code(::Periodic) = Int32(4)
bc_code(::Reflect) = Int32(5)
bc_code(::InPlace) = Int32(6)
bc_code(::InPlaceQ) = Int32(7)
grid_code(::OnGrid) = Int32(0)
grid_code(::OnCell) = Int32(1)
grid_code(::Nothing) = Int32(0)

axis_codes(::NoInterp) = (Int32(7), Int32(0), Int32(0))
function axis_codes(it::BSpline)
    deg = Interpolations.degree(it)
    return (degree_code(deg), bc_code(deg.bc), grid_code(deg.bc.gt))
end
# the per-axis specs of an interpolant: one `BSpline` for all axes or a tuple (mixed.jl)
axis_specs(it::Tuple, N) = it
axis_specs(it, N) = ntuple(_ -> it, N)

"Julia type of a bound / range parameter -> tag: the kernel reproduces the promotion rules from it."
type_tag(::Type{Float32}) = TAG_F32
type_tag(::Type{Float64}) = TAG_F64
type_tag(::Type{<:Union{Integer,Rational}}) = TAG_EXACT
type_tag(::Type{T}) where {T} = throw(ArgumentError("B200Interpolations: unsupported number type $T"))

unwrap(itp::BSplineInterpolation) = (itp, nothing, nothing)
unwrap(s::ScaledInterpolation) = (s.itp::BSplineInterpolation, s, nothing)
function unwrap(e::Union{Extrapolation,FilledExtrapolation})
    b, s, inner = unwrap(e.itp)
    inner === nothing || throw(ArgumentError("B200Interpolations: nested extrapolation is not on the hot path"))
    return (b, s, e)
end

etp_code(f::BoundaryCondition) = bc_code(f)

"""
    descriptor(itp, Tx) -> Desc

Unpack `extrapolate(scale(interpolate(A, it), ranges...), et)` into the POD the kernels read.  Every number comes from
the reference's own functions: `size(parent(coefs))`, `first.(axes(coefs))`, `itp.parentaxes`, `lbounds/ubounds`
(b-splines.jl:141-158, scaling.jl:60-75), `first/step` of the ranges, the flags of `etp.et`.
"""
function descriptor(itp::AbstractInterpolation, ::Type{Tx}) where {Tx<:Union{Float32,Float64}}
    b, s, e = unwrap(itp)
    N = ndims(b)
    N <= MAX_DIMS || throw(ArgumentError("B200Interpolations: at most $MAX_DIMS dimensions"))
    coefs = b.coefs
    specs = axis_specs(b.it, N)
    codes = map(axis_codes, specs)
    lb, ub = lbounds(b), ubounds(b)
    size4 = pad4(map(Int64, size(devarray(coefs))), Int64(1))
    plen = pad4(map(a -> Int64(length(a)), b.parentaxes), Int64(1))
    cfirst = pad4(map(a -> Int32(first(a)), axes(coefs)), Int32(1))
    pfirst = pad4(map(a -> Int32(first(a)), b.parentaxes), Int32(1))
    has_scale = s === nothing ? Int32(0) : Int32(1)
    rtag = rfirst = rstep = olb = oub = otag = nothing
    if s !== nothing
        slb, sub = lbounds(s), ubounds(s)
        rtag = pad4(map(r -> type_tag(eltype(r)), s.ranges), TAG_EXACT)
        rfirst = pad4(map(r -> Float64(first(r)), s.ranges), 0.0)
        rstep = pad4(map(r -> Float64(step(r)), s.ranges), 1.0)
        olb, oub = pad4(map(Float64, slb), 0.0), pad4(map(Float64, sub), 0.0)
        otag = pad4(map(x -> type_tag(typeof(x)), slb), TAG_EXACT)
    else
        rtag, rfirst, rstep = pad4((), TAG_EXACT), pad4((), 0.0), pad4((), 1.0)
        olb, oub, otag = pad4((), 0.0), pad4((), 0.0), pad4((), TAG_EXACT)
    end
    kind, lo, hi, mask = ETP_NONE, pad4((), Int32(0)), pad4((), Int32(0)), Int32(0)
    fill_tag, fill_value = TAG_EXACT, 0.0
    if e isa Extrapolation
        kind = ETP_FLAGS
        et = e.et isa Tuple ? e.et : ntuple(_ -> e.et, N)        # one flag for all axes, or per axis / per side
        lo = pad4(map(f -> f isa Tuple ? etp_code(f[1]) : etp_code(f), et), Int32(0))
        hi = pad4(map(f -> f isa Tuple ? etp_code(f[2]) : etp_code(f), et), Int32(0))
        for (dd, f) in enumerate(et)
            f isa Tuple && (mask |= Int32(1) << (dd - 1))
        end
    elseif e isa FilledExtrapolation
        kind = ETP_FILL
        fill_tag = type_tag(typeof(e.fillvalue))
        fill_value = Float64(e.fillvalue)
    end
    return Desc(Int32(N), dtype_code(eltype(coefs)), dtype_code(Tx), mask,
                size4, plen, cfirst,
                pad4(map(c -> c[1], codes), Int32(1)), pad4(map(c -> c[2], codes), Int32(0)),
                pad4(map(c -> c[3], codes), Int32(0)),
                pad4(map(Float64, lb), 0.0), pad4(map(Float64, ub), 0.0),
                pad4(map(x -> type_tag(typeof(x)), lb), TAG_EXACT),
                has_scale, rtag, rfirst, rstep, olb, oub, otag,
                kind, lo, hi, fill_tag, fill_value, pfirst)
end

coefptr(itp) = devarray(unwrap(itp)[1].coefs).ptr
devof(itp) = devarray(unwrap(itp)[1].coefs).dev
interp_axes(d::Desc) = count(i -> d.degree[i] != 7, 1:d.ndims)   # NoInterp axes carry no gradient component

out_eltype(d::Desc) = ccall((:b200itp_out_dtype, libb200itp), Cint, (Ref{Desc},), Ref(d)) == F32 ? Float32 : Float64

# ------------------------------------------------------------------------------------------------ evaluation seam
struct B200Style{N} <: Base.Broadcast.AbstractArrayStyle{N} end
B200Style{N}(::Val{M}) where {N,M} = B200Style{M}()
Base.BroadcastStyle(::Type{<:B200Array{T,N}}) where {T,N} = B200Style{N}()
Base.BroadcastStyle(a::B200Style{N}, ::Base.Broadcast.DefaultArrayStyle{M}) where {N,M} = B200Style{max(N, M)}()

coord_type(args) = all(a -> eltype(a) === Float32, args) ? Float32 : Float64

"SoA device coordinate arrays of the broadcast shape, element type Tx."
function device_coords(args, shape::Tuple, ::Type{Tx}, dev) where {Tx}
    map(args) do a
        if a isa B200Array{Tx} && size(a) == shape
            a
        else
            h = Array{Tx}(undef, shape)
            h .= (a isa B200Array ? Array(a) : a)      # Julia's own broadcasting: ranges, adjoints, scalars
            B200Array(h; dev=dev)
        end
    end
end

# scratch of the ordered pipeline (about 32 bytes per query), kept and grown per device
const SCRATCH = Dict{Cint,B200Array{UInt8,1}}()
function scratch!(dev::Cint, nbytes::Integer)
    s = get(SCRATCH, dev, nothing)
    if s === nothing || length(s) < nbytes
        s = B200Array{UInt8,1}(undef, (Int(nbytes),); dev=dev)
        SCRATCH[dev] = s
    end
    return s
end

throw_oob(itp, xs, q) = throw(BoundsError(itp, map(x -> Array(x)[q + 1], xs)))   # indexing.jl:6, extrapolation.jl:115-129

"""
`itp.(xs...)`: large scattered broadcasts over arrays beyond the L2 take the ordered pipeline (`b200itp_eval_sorted`,
bit-identical to the direct kernel, which it falls back to by itself outside its scope).
"""
function eval_values(itp::AbstractInterpolation, args, shape)
    Tx = coord_type(args)
    dev = devof(itp)
    xs = device_coords(args, shape, Tx, dev)
    d = descriptor(itp, Tx)
    To = out_eltype(d)
    nq = prod(shape)
    out = B200Array{To,length(shape)}(undef, shape; dev=dev)
    nq == 0 && return out
    ptrs = Ptr{Cvoid}[x.ptr for x in xs]
    foob = Ref{Int64}(-1)
    sbytes = ccall((:b200itp_eval_sorted_scratch_bytes, libb200itp), Csize_t, (Ref{Desc}, Int64), Ref(d), nq)
    GC.@preserve xs ptrs begin
        if sbytes > 0
            s = scratch!(dev, sbytes)
            check(ccall((:b200itp_eval_sorted, libb200itp), Cint,
                        (Cint, Ref{Desc}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Csize_t,
                         Ptr{Cvoid}),
                        dev, Ref(d), coefptr(itp), ptrs, nq, out.ptr, foob, s.ptr, sbytes, C_NULL))
        else
            check(ccall((:b200itp_eval, libb200itp), Cint,
                        (Cint, Ref{Desc}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64},
                         Ptr{Int32}, Ptr{UInt32}, Ptr{Cvoid}),
                        dev, Ref(d), coefptr(itp), ptrs, nq, out.ptr, C_NULL, foob, C_NULL, C_NULL, C_NULL))
        end
    end
    foob[] >= 0 && throw_oob(itp, xs, foob[])
    return out
end

"`Interpolations.gradient.(Ref(itp), xs...)`: an array of `SVector{G,To}`, G = number of interpolating axes."
function eval_gradients(itp::AbstractInterpolation, args, shape)
    Tx = coord_type(args)
    dev = devof(itp)
    xs = device_coords(args, shape, Tx, dev)
    d = descriptor(itp, Tx)
    To = out_eltype(d)
    G = interp_axes(d)
    nq = prod(shape)
    out = B200Array{SVector{G,To},length(shape)}(undef, shape; dev=dev)
    nq == 0 && return out
    ptrs = Ptr{Cvoid}[x.ptr for x in xs]
    foob = Ref{Int64}(-1)
    sbytes = ccall((:b200itp_eval_sorted_scratch_bytes, libb200itp), Csize_t, (Ref{Desc}, Int64), Ref(d), nq)
    GC.@preserve xs ptrs begin
        if sbytes > 0
            s = scratch!(dev, sbytes)
            check(ccall((:b200itp_gradient_sorted, libb200itp), Cint,
                        (Cint, Ref{Desc}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}, Csize_t,
                         Ptr{Cvoid}),
                        dev, Ref(d), coefptr(itp), ptrs, nq, out.ptr, foob, s.ptr, sbytes, C_NULL))
        else
            check(ccall((:b200itp_eval, libb200itp), Cint,
                        (Cint, Ref{Desc}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64},
                         Ptr{Int32}, Ptr{UInt32}, Ptr{Cvoid}),
                        dev, Ref(d), coefptr(itp), ptrs, nq, C_NULL, out.ptr, foob, C_NULL, C_NULL, C_NULL))
        end
    end
    foob[] >= 0 && throw_oob(itp, xs, foob[])
    return out
end

"`Interpolations.hessian.(Ref(itp), xs...)`: an array of `SMatrix{N,N,To}` (symmatrix, indexing.jl:201-229)."
function eval_hessians(itp::AbstractInterpolation, args, shape)
    Tx = coord_type(args)
    dev = devof(itp)
    xs = device_coords(args, shape, Tx, dev)
    d = descriptor(itp, Tx)
    To = out_eltype(d)
    N = Int(d.ndims)
    nq = prod(shape)
    out = B200Array{SMatrix{N,N,To,N * N},length(shape)}(undef, shape; dev=dev)
    nq == 0 && return out
    ptrs = Ptr{Cvoid}[x.ptr for x in xs]
    foob = Ref{Int64}(-1)
    GC.@preserve xs ptrs check(ccall((:b200itp_hessian, libb200itp), Cint,
                                     (Cint, Ref{Desc}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Cvoid}, Ptr{Int64}, Ptr{Cvoid}),
                                     dev, Ref(d), coefptr(itp), ptrs, nq, out.ptr, foob, C_NULL))
    foob[] >= 0 && throw_oob(itp, xs, foob[])
    return out
end

# `itp.(xs...)` reaches `broadcasted(itp, args...)` -> `combine_styles(Ref(itp), args...)` -> B200Style for interpolants
# whose root storage is a B200Array (src/gpu_support.jl:52-57,65-80).  `copy` of the lazy Broadcasted is where Julia's
# generic GPU broadcast would compile a kernel; here it pattern-matches the three calls of the hot path.
function Base.copy(bc::Base.Broadcast.Broadcasted{<:B200Style})
    shape = map(length, axes(bc))
    f = bc.f
    if f isa AbstractInterpolation
        return eval_values(f, bc.args, shape)
    elseif f === Interpolations.gradient && bc.args[1] isa Base.RefValue{<:AbstractInterpolation}
        return eval_gradients(bc.args[1][], Base.tail(bc.args), shape)
    elseif f === Interpolations.hessian && bc.args[1] isa Base.RefValue{<:AbstractInterpolation}
        return eval_hessians(bc.args[1][], Base.tail(bc.args), shape)
    end
    # anything else (e.g. `2 .* jlarray`): host round trip, not on the hot path
    hostargs = map(a -> a isa B200Array ? Array(a) : a, bc.args)
    return B200Array(collect(Base.Broadcast.broadcasted(f, hostargs...)))
end
function Base.copyto!(dest::B200Array, bc::Base.Broadcast.Broadcasted{<:B200Style})
    src = copy(bc)
    size(dest) == size(src) || throw(DimensionMismatch("broadcast destination has the wrong size"))
    return copyto!(dest, src isa B200Array ? src : B200Array(src))
end

# `itp(xvec, yvec, ...)`: tensor-product evaluation (indexing.jl:16-23, scaling.jl:98-100, extrapolation.jl:59-70)
"""
    evalgrid(itp, xvec, yvec, ...) -> B200Array

`itp(xvec, yvec, ...)` of the reference on the device: element `[i, j, ...]` is `itp(xvec[i], yvec[j], ...)`.
"""
function evalgrid(itp::AbstractInterpolation, vecs::AbstractVector...)
    Tx = coord_type(vecs)
    dev = devof(itp)
    d = descriptor(itp, Tx)
    length(vecs) == d.ndims || throw(ArgumentError("evalgrid: need one coordinate vector per dimension"))
    To = out_eltype(d)
    lens = Int64[length(v) for v in vecs]
    xs = map(v -> v isa B200Array{Tx,1} ? v : B200Array(collect(Tx, v isa B200Array ? Array(v) : v); dev=dev), vecs)
    out = B200Array{To,length(vecs)}(undef, Tuple(Int.(lens)); dev=dev)
    prod(lens) == 0 && return out
    ptrs = Ptr{Cvoid}[x.ptr for x in xs]
    foob = Ref{Int64}(-1)
    sbytes = ccall((:b200itp_eval_grid_scratch_bytes, libb200itp), Csize_t, (Ref{Desc}, Ptr{Int64}), Ref(d), lens)
    s = scratch!(dev, max(sbytes, 1))
    GC.@preserve xs ptrs lens check(ccall((:b200itp_eval_grid, libb200itp), Cint,
                                          (Cint, Ref{Desc}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Cvoid}, Ptr{Int64},
                                           Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
                                          dev, Ref(d), coefptr(itp), ptrs, lens, out.ptr, foob, s.ptr, sbytes, C_NULL))
    if foob[] >= 0
        I = CartesianIndices(Tuple(Int.(lens)))[foob[] + 1]
        throw(BoundsError(itp, ntuple(k -> Array(xs[k])[I[k]], length(vecs))))
    end
    return out
end
(itp::BSplineInterpolation{T,N,<:DeviceCoefs})(x::Vararg{AbstractVector,N}) where {T,N} = evalgrid(itp, x...)

# ------------------------------------------------------------------------------------------------ several B200s
# One Julia process drives the devices of one box (SURVEY §8e).  Evaluation needs no collective: queries are
# independent, every device evaluates a contiguous chunk of the query arrays against its replica of the coefficients.

"""
    Comm(devs)

NCCL communicator over the devices `devs` (0-based CUDA ordinals) of this box, driven by this process
(`b200itp_comm_init`: `libnccl.so.2` is loaded at run time, `ncclCommInitAll`).  `close(comm)` destroys it.
"""
mutable struct Comm
    ptr::Ptr{Cvoid}
    devs::Vector{Cint}
    function Comm(devs::AbstractVector{<:Integer})
        d = Cint.(collect(devs))
        p = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve d check(ccall((:b200itp_comm_init, libb200itp), Cint, (Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}),
                                   length(d), d, p))
        c = new(p[], d)
        finalizer(close, c)
        return c
    end
end
function Base.close(c::Comm)
    if c.ptr != C_NULL
        ccall((:b200itp_comm_destroy, libb200itp), Cint, (Ptr{Cvoid},), c.ptr)
        c.ptr = C_NULL
    end
    return nothing
end
Base.length(c::Comm) = length(c.devs)

# the same interpolant around another coefficient buffer (what the reference's adapt_structure does, gpu_support.jl:4-8)
rewrap(itp::BSplineInterpolation{T,N}, buf::B200Array) where {T,N} =
    BSplineInterpolation{T,N}(itp.coefs isa OffsetArray ? OffsetArray(buf, itp.coefs.offsets) : buf, itp.parentaxes, itp.it)

"""
    replicate(comm, itp; root=1) -> Vector

`itp` (coefficients on `comm.devs[root]`) on every device of the communicator, element `i` on `comm.devs[i]`: one NCCL
broadcast of the coefficient array (`b200itp_bcast_coefs`; 512 MiB cross NVLink once).
"""
function replicate(comm::Comm, itp::BSplineInterpolation{T,N,<:DeviceCoefs}; root::Integer=1) where {T,N}
    src = devarray(itp.coefs)
    src.dev == comm.devs[root] || throw(ArgumentError("replicate: the coefficients live on device $(src.dev), not on comm.devs[$root]"))
    bufs = [i == root ? src : B200Array{eltype(src),N}(undef, size(src); dev=comm.devs[i]) for i in 1:length(comm)]
    ptrs = Ptr{Cvoid}[b.ptr for b in bufs]
    GC.@preserve bufs ptrs check(ccall((:b200itp_bcast_coefs, libb200itp), Cint,
                                       (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Csize_t),
                                       comm.ptr, root - 1, ptrs, sizeof(eltype(src)) * length(src)))
    return [rewrap(itp, b) for b in bufs]
end

"Contiguous, order-preserving split of `1:n` into `parts` ranges whose lengths differ by at most one."
function shard_ranges(n::Integer, parts::Integer)
    base, rem = divrem(n, parts)
    los = [(r - 1) * base + min(r - 1, rem) for r in 1:parts]
    return [(los[r] + 1):(los[r] + base + (r <= rem ? 1 : 0)) for r in 1:parts]
end

"""
    interpolate_sharded(comm, A::Array{T,3}, it::BSpline) -> Vector

`interpolate(A, it)` for a large 3-D array with the prefilter sharded over the devices of `comm`
(`b200itp_prefilter_sharded`: z-slabs, local x / y passes, one all-to-all, local z pass, all-gather) — same pass order
and systems as the reference (prefiltering.jl:49-59), same bits as the single-device prefilter.  Returns one
interpolant per device, each with the full coefficient array, ready for `evaluate_sharded`.  Boundary conditions that
pad the array (`padded_axes != axes(A)`) are not sharded: use `replicate(comm, interpolate(b200(A), it))`.
"""
function interpolate_sharded(comm::Comm, A::Array{T,3}, it::BSpline) where {T<:Union{Float32,Float64}}
    Interpolations.padded_axes(axes(A), it) == axes(A) ||
        throw(ArgumentError("interpolate_sharded: $it pads the array; prefilter on one device and `replicate`"))
    G = length(comm)
    zs = shard_ranges(size(A, 3), G)
    slabs = [B200Array(A[:, :, zs[i]]; dev=comm.devs[i]) for i in 1:G]
    full = [B200Array{T,3}(undef, size(A); dev=comm.devs[i]) for i in 1:G]
    sp, fp = Ptr{Cvoid}[s.ptr for s in slabs], Ptr{Cvoid}[f.ptr for f in full]
    codes = axis_codes(it)
    deg, bcs, gts = fill(codes[1], 3), fill(codes[2], 3), fill(codes[3], 3)
    sz = Int64[size(A)...]
    ms = zeros(Cdouble, 2)
    GC.@preserve slabs full sp fp check(ccall((:b200itp_prefilter_sharded, libb200itp), Cint,
                                              (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Ptr{Int32},
                                               Ptr{Int32}, Ptr{Int32}, Ptr{Cdouble}),
                                              comm.ptr, sp, fp, dtype_code(T), sz, deg, bcs, gts, ms))
    TW = Interpolations.tweight(A)
    Tout = typeof(zero(TW) * zero(T))
    return [BSplineInterpolation{Tout,3}(f, axes(A), it) for f in full]
end

"""
    evaluate_sharded(itps, xs...) -> Vector

`itp.(xs...)` with the queries sharded over the replicas `itps` (one per device, from `replicate` or
`interpolate_sharded`): replica `i` evaluates the `i`-th contiguous chunk of the coordinate vectors on its own task;
the result is the concatenation in order, i.e. exactly `itps[1].(xs...)`.
"""
function evaluate_sharded(itps::AbstractVector, xs::AbstractVector...)
    n = length(first(xs))
    all(x -> length(x) == n, xs) || throw(DimensionMismatch("evaluate_sharded: coordinate vectors of different lengths"))
    rs = shard_ranges(n, length(itps))
    tasks = [Threads.@spawn Array(itps[i].(map(x -> view(x, rs[i]), xs)...)) for i in eachindex(itps)]
    return reduce(vcat, map(fetch, tasks))
end

end # module
