# WignerD.jl shim -- the reference package's API over libwignerd_b200.so (B200 engine) via ccall.
#
# Drop-in for jishnub/WignerD.jl v0.1.4: same module name, same exported and qualified functions
# (wignerd, wignerd!, wignerD, wignerD!, WignerD.wignerdjmn, WignerD.wignerDjmn), the same special
# co-latitude types, integer and half-integer j, plus batched array methods that keep the data on one call.
# There is no CUDA.jl, no CPU fallback: every numeric result comes from the sm_100a kernels behind the C ABI
# declared in include/wignerd_b200.h.
#
# NOT EXECUTED in this repository's CI: Julia is not present in the build image (see DESIGN.md); the
# identical binding is exercised from Python (wignerd/jl_b200/api.py, ctypes) by tests/test_gpu_parity.py.
module WignerD

using HalfIntegers

export wignerd!, wignerd, wignerD!, wignerD

const libwd = get(ENV, "WIGNERD_B200_LIB", "libwignerd_b200")

# ---- status codes of include/wignerd_b200.h ---------------------------------------------------------------
const WD_OK, WD_ERR_ARGUMENT, WD_ERR_ASSERT, WD_ERR_CUDA, WD_ERR_NOMEM, WD_ERR_UNSUPPORTED = 0:5
const WD_MEM_HOST = Cint(0)

function check(rc::Cint)
    rc == WD_OK && return nothing
    msg = unsafe_string(ccall((:wd_last_error, libwd), Cstring, ()))
    rc == WD_ERR_ARGUMENT && throw(ArgumentError(msg))      # src/WignerD.jl:191-194 of the reference
    rc == WD_ERR_ASSERT && throw(AssertionError(msg))       # :96-97, :148-149, :267
    rc == WD_ERR_NOMEM && throw(OutOfMemoryError())
    error("libwignerd_b200: " * msg)
end

# ---- one handle (GPU, stream, per-j table cache) per Task, like the reference's TaskLocalValue cache ------
mutable struct Handle
    ptr::Ptr{Cvoid}
end
function Handle(device::Integer = parse(Int, get(ENV, "WIGNERD_B200_DEVICE", "0")))
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:wd_create, libwd), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ref))
    h = Handle(ref[])
    finalizer(x -> ccall((:wd_destroy, libwd), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end
# several GPUs behind one handle (wd_create_multi): the batched methods then shard their host arrays over the devices
function Handle(devices::AbstractVector{<:Integer})
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    devs = Vector{Cint}(devices)
    check(ccall((:wd_create_multi, libwd), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), length(devs), devs, ref))
    h = Handle(ref[])
    finalizer(x -> ccall((:wd_destroy, libwd), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end
handle() = get!(() -> Handle(), task_local_storage(), :wignerd_b200_handle)::Handle
# use_devices!([0, 1, ..., 7]): this Task's calls run on all listed GPUs from now on
use_devices!(devices::AbstractVector{<:Integer}) = (task_local_storage(:wignerd_b200_handle, Handle(devices)); nothing)
ndevices() = Int(ccall((:wd_multi_size, libwd), Cint, (Ptr{Cvoid},), handle().ptr))

# persisted eigenvector tables: a cold process loads them instead of running the eigensolver (extends Jy_eigen's cache)
save_tables(path::AbstractString) = check(ccall((:wd_tables_save, libwd), Cint, (Ptr{Cvoid}, Cstring), handle().ptr, path))
load_tables(path::AbstractString) = check(ccall((:wd_tables_load, libwd), Cint, (Ptr{Cvoid}, Cstring), handle().ptr, path))

# ---- special co-latitudes: Real subtypes accepted wherever beta is ----------------------------------------
abstract type SpecialCoLatitudes <: Real end
struct Equator <: SpecialCoLatitudes end
struct NorthPole <: SpecialCoLatitudes end
struct SouthPole <: SpecialCoLatitudes end
Base.Float64(::NorthPole) = 0.0
Base.Float64(::SouthPole) = Float64(pi)
Base.Float64(::Equator) = Float64(pi) / 2
Base.AbstractFloat(p::SpecialCoLatitudes) = Float64(p)
Base.promote_rule(::Type{<:SpecialCoLatitudes}, T::Type{<:Real}) = promote_type(Float64, T)
Base.cos(::NorthPole) = 1.0;  Base.sin(::NorthPole) = 0.0
Base.cos(::SouthPole) = -1.0; Base.sin(::SouthPole) = 0.0
Base.cos(::Equator) = 0.0;    Base.sin(::Equator) = 1.0
betakind(::Real) = Cint(0)
betakind(::NorthPole) = Cint(1)
betakind(::SouthPole) = Cint(2)
betakind(::Equator) = Cint(3)

# exact (sin, cos) of alpha*pi/2 for integer / half-integer alpha (used by the reference's tests)
function _sincos(alpha::Integer, ::Equator)
    c, s = ((1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0))[mod(Int(alpha), 4) + 1]
    return s, c
end
function _sincos(alpha::HalfInteger, ::Equator)
    r = 1 / sqrt(2)
    c, s = ((1.0, 0.0), (r, r), (0.0, 1.0), (-r, r), (-1.0, 0.0), (-r, -r), (0.0, -1.0), (r, -r))[mod(Int(2alpha), 8) + 1]
    return s, c
end

# ---- j / m plumbing: the C ABI takes 2j, 2m, 2n as Int32 ----------------------------------------------------
twice(x) = Int32(Int(2x))                    # InexactError for anything that is not an integer or half-integer
_matrixsize(j) = (Int(2j + 1), Int(2j + 1))

struct WignerMatrix{T, A <: AbstractMatrix{T}} <: AbstractMatrix{T}
    twoj::Int
    D::A
    function WignerMatrix(j, D::AbstractMatrix{T}) where {T}
        @assert j >= 0 "j must be >= 0"
        @assert size(D) == _matrixsize(j) "size of matrix must be $(Int(2j + 1))x$(Int(2j + 1))"
        new{T, typeof(D)}(Int(2j), D)
    end
end
Base.size(w::WignerMatrix) = size(w.D)
Base.parent(w::WignerMatrix) = w.D
Base.IndexStyle(::Type{<:WignerMatrix{<:Any, A}}) where {A} = IndexStyle(A)
_index(w::WignerMatrix, m) = (Int(2m) + w.twoj) ÷ 2 + 1
Base.getindex(w::WignerMatrix, m::Real, n::Real) = w.D[_index(w, m), _index(w, n)]
Base.setindex!(w::WignerMatrix, v, m::Real, n::Real) = (w.D[_index(w, m), _index(w, n)] = v; w)
_offsetmatrix(j, D::AbstractMatrix) = WignerMatrix(j, D)

# the optional Jy scratch is validated like the reference does on a cache miss, then ignored
function checkJy(j, Jy)
    @assert eltype(Jy) == ComplexF64 "preallocated Jy matrix must have an element type of ComplexF64"
    @assert size(Jy) == _matrixsize(j) "preallocated Jy matrix must be of size $(_matrixsize(j))"
end

# ---- single elements ------------------------------------------------------------------------------------------
function wignerdjmn(j, m, n, beta::Real, Jy...)
    out = Ref{Float64}()
    check(ccall((:wd_djmn_batch, libwd), Cint,
                (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Float64}, Int64, Cint, Ref{Float64}, Cint),
                handle().ptr, twice(j), twice(m), twice(n), Float64(beta), 1, betakind(beta), out, WD_MEM_HOST))
    return out[]
end
function wignerDjmn(j, m, n, alpha::Real, beta::Real, gamma::Real, Jy...)
    out = Ref{ComplexF64}()
    check(ccall((:wd_Djmn_batch, libwd), Cint,
                (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Int64, Cint,
                 Ref{ComplexF64}, Cint),
                handle().ptr, twice(j), twice(m), twice(n), Float64(alpha), Float64(beta), Float64(gamma), 1,
                betakind(beta), out, WD_MEM_HOST))
    return out[]
end

# ---- matrices -------------------------------------------------------------------------------------------------
function wignerd!(d::Matrix{Float64}, j, beta::Real, Jy...)
    @boundscheck @assert size(d) == _matrixsize(j) "size of dj must be $(Int(2j + 1))x$(Int(2j + 1))"
    isempty(Jy) || checkJy(j, Jy[1])
    check(ccall((:wd_d_matrix, libwd), Cint, (Ptr{Cvoid}, Int32, Float64, Cint, Ptr{Float64}),
                handle().ptr, twice(j), Float64(beta), betakind(beta), d))
    return d
end
function wignerd!(d::AbstractMatrix, j, beta::Real, Jy...)        # offset / strided / non-Float64 destinations
    tmp = Matrix{Float64}(d)
    wignerd!(tmp, j, beta, Jy...)
    copyto!(d, tmp)
    return d
end
wignerd(j, beta::Real, Jy...) = wignerd!(zeros(_matrixsize(j)), j, beta, Jy...)

function wignerD!(D::Matrix{ComplexF64}, j, alpha::Real, beta::Real, gamma::Real, Jy...)
    @boundscheck @assert size(D) == _matrixsize(j) "size of dj must be $(Int(2j + 1))x$(Int(2j + 1))"
    isempty(Jy) || checkJy(j, Jy[1])
    check(ccall((:wd_D_matrix, libwd), Cint, (Ptr{Cvoid}, Int32, Float64, Float64, Cint, Float64, Ptr{ComplexF64}),
                handle().ptr, twice(j), Float64(alpha), Float64(beta), betakind(beta), Float64(gamma), D))
    return D
end
function wignerD!(D::AbstractMatrix, j, alpha::Real, beta::Real, gamma::Real, Jy...)
    tmp = Matrix{ComplexF64}(D)
    wignerD!(tmp, j, alpha, beta, gamma, Jy...)
    copyto!(D, tmp)
    return D
end
wignerD(j, alpha::Real, beta::Real, gamma::Real, Jy...) =
    wignerD!(zeros(ComplexF64, _matrixsize(j)), j, alpha, beta, gamma, Jy...)

# ---- batched methods (what the GPU is for): N x N x nb arrays, one call ----------------------------------------
function wignerd(j, beta::AbstractVector{<:Real})
    N = Int(2j + 1)
    b = Vector{Float64}(beta)
    out = Array{Float64}(undef, N, N, length(b))
    check(ccall((:wd_d_batch, libwd), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Cint),
                handle().ptr, twice(j), b, length(b), out, WD_MEM_HOST))
    return out
end
function wignerD(j, alpha::AbstractVector{<:Real}, beta::AbstractVector{<:Real}, gamma::AbstractVector{<:Real})
    N = Int(2j + 1)
    a, b, g = Vector{Float64}(alpha), Vector{Float64}(beta), Vector{Float64}(gamma)
    @assert length(a) == length(b) == length(g)
    out = Array{ComplexF64}(undef, N, N, length(b))
    check(ccall((:wd_D_batch, libwd), Cint,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{ComplexF64}, Cint),
                handle().ptr, twice(j), a, b, g, length(b), out, WD_MEM_HOST))
    return out
end
function wignerdjmn(j::AbstractVector, m::AbstractVector, n::AbstractVector, beta::AbstractVector{<:Real})
    tj, tm, tn = Int32.(2 .* j), Int32.(2 .* m), Int32.(2 .* n)
    b = Vector{Float64}(beta)
    out = Vector{Float64}(undef, length(b))
    check(ccall((:wd_djmn_batch, libwd), Cint,
                (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Cint),
                handle().ptr, tj, tm, tn, b, length(b), Cint(0), out, WD_MEM_HOST))
    return out
end


# rotated spherical-harmonic coefficients, out[l^2 + l + m' + 1, b] = sum_m D^l_{m m'}(alpha_b, beta_b, gamma_b) alm[l^2 + l + m + 1 (, b)]
# for l = 0..lmax (docs/src/index.md:76-84 of the reference), without forming a D matrix.  alm: a vector of (lmax+1)^2 coefficients
# rotated by every rotation, or a (lmax+1)^2 x nb matrix (one set per rotation).
function rotate_sh(lmax::Integer, alm::AbstractVecOrMat{<:Number}, alpha::AbstractVector{<:Real}, beta::AbstractVector{<:Real},
                   gamma::AbstractVector{<:Real})
    L2 = (Int(lmax) + 1)^2
    a, b, g = Vector{Float64}(alpha), Vector{Float64}(beta), Vector{Float64}(gamma)
    @assert length(a) == length(b) == length(g)
    A = Array{ComplexF64}(alm)
    @assert size(A, 1) == L2 "alm must hold (lmax+1)^2 coefficients per set"
    stride = ndims(A) == 1 ? 0 : L2
    ndims(A) == 2 && @assert size(A, 2) == length(b)
    out = Array{ComplexF64}(undef, L2, length(b))
    check(ccall((:wd_rotate_sh, libwd), Cint,
                (Ptr{Cvoid}, Int32, Ptr{ComplexF64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{ComplexF64}, Cint),
                handle().ptr, Int32(lmax), A, stride, a, b, g, length(b), out, WD_MEM_HOST))
    return out
end

end # module
