# WaveletsB200.jl -- Julia shim: routes Wavelets.jl's transform dispatch points to libwavelets_b200.so.
#
# UNTESTED IN THIS REPOSITORY'S BUILD IMAGE (no Julia runtime there); it is the binding a maintainer adds on a box that
# has Julia + CUDA.jl.  It defines methods of the SAME four families the reference's KernelAbstractions extension
# overrides (ext/WaveletsGPUExt/WaveletsGPUExt.jl:11) for `CuArray`s, so that `dwt`, `idwt`, `dwt!`, `idwt!`, `wpt`,
# `iwpt`, `wpt!`, `iwpt!` and `wavelet(...)` keep working unchanged, and adds the column-wise `dwtc`/`idwtc` the
# reference only stubs (src/Transforms/transforms_main.jl:179-181).
module WaveletsB200

using Wavelets
using Wavelets.WT: OrthoFilter, GLS, DiscreteWavelet
using CUDA
import Wavelets.Transforms: _dwt!, _wpt!
import Wavelets.Util: maxtransformlevels, maxmodwttransformlevels, isvalidtree, maketree, sufficientpoweroftwo, iscube

const LIB = get(ENV, "WAVELETS_B200_LIB", "libwavelets_b200.so")

# ---- C ABI mirror (include/wavelets_b200.h) ---------------------------------------------------------------------
const WB200_MAX_LIFT_COEF = 8
struct LiftStep                      # wb200_lift_step
    is_predict::Int32
    shift::Int32
    nc::Int32
    coef::NTuple{WB200_MAX_LIFT_COEF, Float64}
end
function LiftStep(s::Wavelets.WT.LSStep)
    c = s.param.coef
    length(c) <= WB200_MAX_LIFT_COEF || throw(ArgumentError("lifting step has too many coefficients"))
    coef = ntuple(i -> i <= length(c) ? Float64(c[i]) : 0.0, WB200_MAX_LIFT_COEF)
    LiftStep(Int32(s.steptype isa Wavelets.WT.PredictStep), Int32(s.param.shift), Int32(length(c)), coef)
end

dtype_code(::Type{Float32}) = Int32(0)
dtype_code(::Type{Float64}) = Int32(1)
dtype_code(::Type{ComplexF32}) = Int32(2)
dtype_code(::Type{ComplexF64}) = Int32(3)

# status -> the reference's exception (transforms_filter.jl:25-34, transforms_lifting.jl:34-39,131-140)
function check(rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:wb200_status_string, LIB), Cstring, (Int32,), rc))
    rc == 1 && throw(DimensionMismatch(msg))
    rc in (2, 3, 4, 5, 6, 8) && throw(ArgumentError(msg))
    detail = unsafe_string(ccall((:wb200_last_error_string, LIB), Cstring, ()))
    error("wavelets_b200: $msg ($detail)")
end

dims3(x) = Int64[size(x)..., ones(Int, 3 - ndims(x))...]
flags() = UInt32(get(ENV, "WAVELETS_B200_STRICT_FP", "0") == "1" ? 1 : 0)

# ---- filter transforms: _dwt!(y, x, filter, L, fw)   (src/Transforms/transforms_filter.jl:13,113,192) ------------
function _dwt!(y::CuArray{T,N}, x::CuArray{T,N}, filter::OrthoFilter, L::Integer, fw::Bool) where {T<:Union{Float32,Float64,ComplexF32,ComplexF64},N}
    size(x) == size(y) || throw(DimensionMismatch("in and out array size must match"))
    dwt_filter!(y, x, N, 1, filter, L, fw)
end
function dwt_filter!(y::CuArray{T}, x::CuArray{T}, nd::Int, batch::Integer, filter::OrthoFilter, L::Integer, fw::Bool) where T
    qmf = Vector{Float64}(filter.qmf)
    d = Int64[size(x)[1:nd]..., ones(Int, 3 - nd)...]
    rc = ccall((:wb200_dwt_filter, LIB), Int32,
               (CuPtr{Cvoid}, CuPtr{Cvoid}, Int32, Ptr{Int64}, Int64, Ptr{Float64}, Int32, Int32, Int32, Int32,
                CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}, UInt32),
               pointer(y), pointer(x), nd, d, batch, qmf, length(qmf), L, fw, dtype_code(T),
               CU_NULL, 0, CUDA.stream().handle, flags())
    check(rc)
    return y
end

# ---- lifting transforms: _dwt!(y, scheme, L, fw)   (src/Transforms/transforms_lifting.jl:30,128,200) -------------
function _dwt!(y::CuArray{T,N}, scheme::GLS, L::Integer, fw::Bool) where {T<:Union{Float32,Float64,ComplexF32,ComplexF64},N}
    dwt_lifting!(y, y, N, 1, scheme, L, fw)          # x == y: in-place form
end
function dwt_lifting!(y::CuArray{T}, x::CuArray{T}, nd::Int, batch::Integer, scheme::GLS, L::Integer, fw::Bool) where T
    steps = LiftStep.(scheme.step)
    d = Int64[size(x)[1:nd]..., ones(Int, 3 - nd)...]
    rc = ccall((:wb200_dwt_lifting, LIB), Int32,
               (CuPtr{Cvoid}, CuPtr{Cvoid}, Int32, Ptr{Int64}, Int64, Ptr{LiftStep}, Int32, Float64, Float64, Int32, Int32, Int32,
                CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}, UInt32),
               pointer(y), pointer(x), nd, d, batch, steps, length(steps), scheme.norm1, scheme.norm2, L, fw, dtype_code(T),
               CU_NULL, 0, CUDA.stream().handle, flags())
    check(rc)
    return y
end
# allocating forms skip the reference's copyto! (transforms_main.jl:119-124): out-of-place call, x != y
Wavelets.dwt(x::CuArray{T}, scheme::GLS, L::Integer=maxtransformlevels(x)) where T = dwt_lifting!(similar(x), x, ndims(x), 1, scheme, L, true)
Wavelets.idwt(x::CuArray{T}, scheme::GLS, L::Integer=maxtransformlevels(x)) where T = dwt_lifting!(similar(x), x, ndims(x), 1, scheme, L, false)

# ---- wavelet packets: _wpt!(y, x, filter, tree, fw) / _wpt!(y, scheme, tree, fw) --------------------------------
function _wpt!(y::CuVector{T}, x::CuVector{T}, filter::OrthoFilter, tree::BitVector, fw::Bool) where T
    qmf = Vector{Float64}(filter.qmf); t = Vector{UInt8}(tree)
    rc = ccall((:wb200_wpt_filter, LIB), Int32,
               (CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Ptr{Float64}, Int32, Ptr{UInt8}, Int64, Int32, Int32,
                CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}, UInt32),
               pointer(y), pointer(x), length(x), 1, qmf, length(qmf), t, length(t), fw, dtype_code(T),
               CU_NULL, 0, CUDA.stream().handle, flags())
    check(rc); y
end
function _wpt!(y::CuVector{T}, scheme::GLS, tree::BitVector, fw::Bool) where T
    steps = LiftStep.(scheme.step); t = Vector{UInt8}(tree)
    rc = ccall((:wb200_wpt_lifting, LIB), Int32,
               (CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Ptr{LiftStep}, Int32, Float64, Float64, Ptr{UInt8}, Int64, Int32, Int32,
                CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}, UInt32),
               pointer(y), pointer(y), length(y), 1, steps, length(steps), scheme.norm1, scheme.norm2, t, length(t), fw, dtype_code(T),
               CU_NULL, 0, CUDA.stream().handle, flags())
    check(rc); y
end

# ---- maximal-overlap DWT: modwt(x, wt, L) / imodwt(xw, wt)  (transforms_maximal_overlap.jl:44-62, 98-108) ----------
function Wavelets.Transforms.modwt(x::CuVector{T}, wt::OrthoFilter, L::Integer=maxmodwttransformlevels(x)) where {T<:Union{Float32,Float64}}
    L <= maxmodwttransformlevels(x) || throw(ArgumentError("Too many transform levels (length(x) < 2^L)"))
    L >= 1 || throw(ArgumentError("L must be >= 1"))
    qmf = Vector{Float64}(wt.qmf); y = CuArray{T}(undef, length(x), L + 1)
    rc = ccall((:wb200_modwt, LIB), Int32,
               (CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Ptr{Float64}, Int32, Int32, Int32, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}, UInt32),
               pointer(y), pointer(x), length(x), 1, qmf, length(qmf), L, dtype_code(T), CU_NULL, 0, CUDA.stream().handle, flags())
    check(rc); y
end
function Wavelets.Transforms.imodwt(xw::CuMatrix{T}, wt::OrthoFilter) where {T<:Union{Float32,Float64}}
    qmf = Vector{Float64}(wt.qmf); x = CuArray{T}(undef, size(xw, 1))
    rc = ccall((:wb200_imodwt, LIB), Int32,
               (CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, Ptr{Float64}, Int32, Int32, Int32, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}, UInt32),
               pointer(x), pointer(xw), size(xw, 1), 1, qmf, length(qmf), size(xw, 2), dtype_code(T), CU_NULL, 0, CUDA.stream().handle, flags())
    check(rc); x
end

# ---- Threshold: threshold!, noisest, denoise on the device (src/Threshold/threshold_main.jl, denoising.jl) ---------------
th_code(::Wavelets.Threshold.HardTH) = Int32(0);     th_code(::Wavelets.Threshold.SoftTH) = Int32(1)
th_code(::Wavelets.Threshold.SemiSoftTH) = Int32(2); th_code(::Wavelets.Threshold.SteinTH) = Int32(3)
th_code(::Wavelets.Threshold.NegTH) = Int32(4);      th_code(::Wavelets.Threshold.PosTH) = Int32(5)
function Wavelets.Threshold.threshold!(x::CuArray{T}, TH::Wavelets.Threshold.THType, t::Real=0) where {T<:Union{Float32,Float64}}
    rc = ccall((:wb200_threshold, LIB), Int32, (CuPtr{Cvoid}, Int64, Int32, Float64, Int32, Ptr{Cvoid}),
               pointer(x), length(x), th_code(TH), Float64(t), dtype_code(T), CUDA.stream().handle)
    check(rc); x
end
function Wavelets.Threshold.threshold!(x::CuArray{T}, ::Wavelets.Threshold.BiggestTH, m::Int) where {T<:Union{Float32,Float64}}
    rc = ccall((:wb200_threshold_biggest, LIB), Int32, (CuPtr{Cvoid}, Int64, Int64, Int32, Ptr{Cvoid}),
               pointer(x), length(x), m, dtype_code(T), CUDA.stream().handle)
    check(rc); x
end
# wkind / qmf / steps of a wavelet (nothing, OrthoFilter, GLS)
wt_args(::Nothing) = (Int32(0), Float64[], LiftStep[], 0.0, 0.0)
wt_args(f::OrthoFilter) = (Int32(1), Vector{Float64}(f.qmf), LiftStep[], 0.0, 0.0)
wt_args(s::GLS) = (Int32(2), Float64[], LiftStep.(s.step), s.norm1, s.norm2)
function Wavelets.Threshold.noisest(x::CuArray{T,N}, wt::Union{DiscreteWavelet,Nothing}=Wavelets.Threshold.DEFAULT_WAVELET, L::Integer=1) where {T<:Union{Float32,Float64},N}
    wk, qmf, steps, n1, n2 = wt_args(wt); out = Ref{Float64}(0)
    rc = ccall((:wb200_noisest, LIB), Int32,
               (Ptr{Float64}, CuPtr{Cvoid}, Int32, Ptr{Int64}, Int32, Ptr{Float64}, Int32, Ptr{LiftStep}, Int32, Float64, Float64, Int32, Int32, Ptr{Cvoid}, UInt32),
               out, pointer(x), N, dims3(x), wk, qmf, length(qmf), steps, length(steps), n1, n2, Int32(L), dtype_code(T), CUDA.stream().handle, flags())
    check(rc); out[]
end
function Wavelets.Threshold.denoise(x::CuArray{T,N}, wt::Union{DiscreteWavelet,Nothing}=Wavelets.Threshold.DEFAULT_WAVELET;
                                    L::Int=min(maxtransformlevels(x), 6), dnt::Wavelets.Threshold.VisuShrink=Wavelets.Threshold.VisuShrink(size(x, 1)),
                                    estnoise::Union{Function,Nothing}=nothing, TI::Bool=false,
                                    nspin::Union{Int,Tuple}=ntuple(_ -> 8, N)) where {T<:Union{Float32,Float64},N}
    wk, qmf, steps, n1, n2 = wt_args(wt)
    sigma = estnoise === nothing ? NaN : Float64(estnoise(x, wt))       # NaN: noisest on the device, no host round trip
    spin = Int32[nspin..., ones(Int32, 3 - length(nspin))...]
    y = similar(x)
    rc = ccall((:wb200_denoise, LIB), Int32,
               (CuPtr{Cvoid}, CuPtr{Cvoid}, Int32, Ptr{Int64}, Int32, Ptr{Float64}, Int32, Ptr{LiftStep}, Int32, Float64, Float64, Int32,
                Int32, Float64, Float64, Int32, Ptr{Int32}, Int32, Ptr{Cvoid}, UInt32),
               pointer(y), pointer(x), N, dims3(x), wk, qmf, length(qmf), steps, length(steps), n1, n2, L,
               th_code(dnt.th), dnt.t, sigma, TI, spin, dtype_code(T), CUDA.stream().handle, flags())
    check(rc); y
end

# ---- best basis: coefentropy / bestbasistree (src/Threshold/entropy.jl) -------------------------------------------------
et_code(::Wavelets.Threshold.ShannonEntropy) = Int32(0); et_code(::Wavelets.Threshold.LogEnergyEntropy) = Int32(1)
function Wavelets.Threshold.coefentropy(x::CuArray{T}, et::Wavelets.Threshold.Entropy, nrm::T=T(NaN)) where {T<:Union{Float32,Float64}}
    out = Ref{Float64}(0)
    rc = ccall((:wb200_coefentropy, LIB), Int32, (Ptr{Float64}, CuPtr{Cvoid}, Int64, Int32, Float64, Int32, Ptr{Cvoid}),
               out, pointer(x), length(x), et_code(et), Float64(nrm), dtype_code(T), CUDA.stream().handle)
    check(rc); T(out[])
end
function Wavelets.Threshold.bestbasistree(y::CuVector{T}, wt::DiscreteWavelet, tree::BitVector,
                                          et::Wavelets.Threshold.Entropy=Wavelets.Threshold.ShannonEntropy()) where {T<:Union{Float32,Float64}}
    wk, qmf, steps, n1, n2 = wt_args(wt); t = Vector{UInt8}(tree); best = similar(t)
    rc = ccall((:wb200_bestbasistree, LIB), Int32,
               (Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, CuPtr{Cvoid}, Int64, Int32, Ptr{Float64}, Int32, Ptr{LiftStep}, Int32, Float64, Float64,
                Ptr{UInt8}, Int64, Int32, Int32, Ptr{Cvoid}, UInt32),
               best, C_NULL, C_NULL, pointer(y), length(y), wk, qmf, length(qmf), steps, length(steps), n1, n2, t, length(t), et_code(et),
               dtype_code(T), CUDA.stream().handle, flags())
    check(rc); BitVector(best .!= 0)
end
Wavelets.Threshold.bestbasistree(y::CuVector{T}, wt::DiscreteWavelet, L::Integer=maxtransformlevels(y),
                                 et::Wavelets.Threshold.Entropy=Wavelets.Threshold.ShannonEntropy()) where {T<:Union{Float32,Float64}} =
    Wavelets.Threshold.bestbasistree(y, wt, maketree(length(y), L, :full), et)

# ---- column-wise batch forms (the last dimension indexes independent signals / images) ---------------------------
export dwtc, idwtc
dwtc(x::CuArray, wt::OrthoFilter, L::Integer=minimum(maxtransformlevels.(size(x)[1:end-1]))) =
    dwt_filter!(similar(x), x, ndims(x) - 1, size(x)[end], wt, L, true)
idwtc(x::CuArray, wt::OrthoFilter, L::Integer=minimum(maxtransformlevels.(size(x)[1:end-1]))) =
    dwt_filter!(similar(x), x, ndims(x) - 1, size(x)[end], wt, L, false)
dwtc(x::CuArray, wt::GLS, L::Integer=minimum(maxtransformlevels.(size(x)[1:end-1]))) =
    dwt_lifting!(similar(x), x, ndims(x) - 1, size(x)[end], wt, L, true)
idwtc(x::CuArray, wt::GLS, L::Integer=minimum(maxtransformlevels.(size(x)[1:end-1]))) =
    dwt_lifting!(similar(x), x, ndims(x) - 1, size(x)[end], wt, L, false)

end # module
