# PDMPFluxCUDA.jl -- Julia `ccall` glue over libpdmpflux_cuda.so (include/pdmpflux_cuda.h).
#
# UNTESTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The tested binding of the same C ABI is the
# ctypes one in pdmpflux.jl_b200/_lib.py; this file is what a PDMPFlux.jl maintainer would add (see
# INTEGRATION.md).  It returns real `PDMPFlux.PDMPHistory` objects, so `diagnostic`, `plot_traj`, `anim_traj`
# and `sample_from_skeleton` of the reference work unchanged on the result.
module PDMPFluxCUDA

using PDMPFlux: PDMPHistory
export CuPotential, GaussStd, GaussDiag, GaussEquicorr, Banana, BananaReadmeScalar,
       CuZigZag, CuBPS, CuForwardECMC, CuBoomerang, sample_skeleton, sample_from_skeleton, sample

const LIB = get(ENV, "PDMPFLUX_CUDA_LIB", joinpath(@__DIR__, "..", "pdmpflux.jl_b200", "lib", "libpdmpflux_cuda.so"))

# ---- C structs (layout must match include/pdmpflux_cuda.h) -----------------------------------------------------
struct CConfig
    grid_size::Int32; vectorized_bound::Int32; signed_bound::Int32; adaptive::Int32; deriv_mode::Int32
    gaussian_velocity::Int32; ran_p::Int32; switch_::Int32; positive::Int32; max_steps::Int32
    tmax::Float64; refresh_rate::Float64; mix_p::Float64; speed_factor::Float64
end
struct CHistory
    X::Ptr{Float64}; V::Ptr{Float64}; t::Ptr{Float64}; horizon::Ptr{Float64}; ar::Ptr{Float64}
    error_value_ar::Ptr{Float64}; errored_bound::Ptr{Int32}; rejected::Ptr{Int32}; hitting_horizon::Ptr{Int32}
    status::Ptr{Int32}; tape_pos::Ptr{Int64}; counters::Ptr{Int64}; n_cols::Int64; on_device::Int32
end

last_error() = unsafe_string(ccall((:pdmpflux_last_error, LIB), Cstring, ()))
function check(rc::Cint)
    rc == 0 && return
    msg = last_error()
    rc == -1 && throw(ArgumentError(msg))          # PDMPFLUX_ERR_ARGUMENT
    rc == -2 && throw(DimensionMismatch(msg))      # PDMPFLUX_ERR_DIMENSION_MISMATCH
    error("libpdmpflux_cuda ($rc): $msg")          # unsupported / CUDA / chain failure: no CPU fallback
end

# ---- device potentials: replace the `∇U` closure -----------------------------------------------------------------
struct CuPotential
    kind::Int32
    params::Vector{Float64}
end
GaussStd() = CuPotential(0, Float64[])
GaussDiag(p::AbstractVector) = CuPotential(1, collect(Float64, p))
GaussEquicorr(rho::Real) = CuPotential(2, [Float64(rho)])
Banana() = CuPotential(3, Float64[])
BananaReadmeScalar() = CuPotential(4, Float64[])

mutable struct CuPDMP
    kind::Int32; dim::Int; pot::Ptr{Cvoid}; handle::Ptr{Cvoid}; flow_kind::Int32
    state::Any
end
deriv_mode(ad::String) = ad in ("", "Undefined", "FiniteDiff") ? Int32(1) : Int32(0)

function CuPDMP(kind, dim, pot::CuPotential, cfg::CConfig)
    hp = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pdmpflux_potential_create, LIB), Cint, (Cint, Cint, Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
                pot.kind, dim, pot.params, length(pot.params), hp))
    hs = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pdmpflux_sampler_create, LIB), Cint, (Cint, Cint, Ptr{Cvoid}, Ref{CConfig}, Ref{Ptr{Cvoid}}),
                kind, dim, hp[], Ref(cfg), hs))
    s = CuPDMP(kind, dim, hp[], hs[], kind == 3 ? 1 : 0, nothing)
    finalizer(s) do x
        ccall((:pdmpflux_sampler_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle)
        ccall((:pdmpflux_potential_destroy, LIB), Cint, (Ptr{Cvoid},), x.pot)
    end
    return s
end

# constructor keyword surface of the reference (ZigZagSamplers.jl:58-60, BouncyParticleSamplers.jl:21-24,
# ForwardEventChainMonteCarlo.jl:301-303, BoomerangSamplers.jl:21-23)
CuZigZag(dim::Int, pot::CuPotential; grid_size::Int=10, tmax=2.0, refresh_rate::Float64=0.0, vectorized_bound::Bool=true,
         signed_bound::Bool=true, adaptive::Bool=true, AD_backend::String="FiniteDiff") =
    CuPDMP(0, dim, pot, CConfig(grid_size, vectorized_bound, signed_bound, adaptive, deriv_mode(AD_backend), 0, 0, 1, 1, 0,
                                Float64(tmax), refresh_rate, 0.5, 1.0))
CuBPS(dim::Int, pot::CuPotential; grid_size::Int=10, tmax=1.0, refresh_rate::Float64=0.1, signed_bound::Bool=true,
      adaptive::Bool=true, AD_backend::String="ForwardDiff", Gaussian_velocity::Bool=false) =
    CuPDMP(1, dim, pot, CConfig(grid_size, 0, signed_bound, adaptive, deriv_mode(AD_backend), Gaussian_velocity, 0, 1, 1, 0,
                                Float64(tmax), refresh_rate, 0.5, 1.0))
CuForwardECMC(dim::Int, pot::CuPotential; grid_size::Int=10, tmax=2.0, signed_bound::Bool=true, adaptive::Bool=true,
              ran_p::Bool=false, mix_p::Float64=0.5, switch::Bool=true, positive::Bool=true,
              AD_backend::String="ForwardDiff", speed_factor::Float64=1.0) =
    CuPDMP(2, dim, pot, CConfig(grid_size, 0, signed_bound, adaptive, deriv_mode(AD_backend), 0, ran_p, switch, positive, 0,
                                Float64(tmax), 0.0, mix_p, speed_factor))
CuBoomerang(dim::Int, pot::CuPotential; grid_size::Int=10, tmax=1.0, refresh_rate::Float64=0.1, signed_bound::Bool=true,
            adaptive::Bool=true, AD_backend::String="FiniteDiff") =
    CuPDMP(3, dim, pot, CConfig(grid_size, 0, signed_bound, adaptive, deriv_mode(AD_backend), 0, 0, 1, 1, 0,
                                Float64(tmax), refresh_rate, 0.5, 1.0))

"""
    sample_skeleton(sampler::CuPDMP, n_sk, xinit, vinit; seed) -> PDMPHistory            (one chain, as upstream)
    sample_skeleton(sampler::CuPDMP, n_sk, xinit::Matrix, vinit::Matrix; seed) -> Vector{PDMPHistory}  (d x C inits)

Replaces `PDMPFlux.sample_skeleton` (src/sample.jl:253-284).  Each chain's slab is written by the library directly
into the `Matrix{Float64}(d, n_sk)` / `Vector` storage of a `PDMPHistory` (chain-major layout, no copies).
"""
function sample_skeleton(s::CuPDMP, n_sk::Int, xinit::Matrix{Float64}, vinit::Matrix{Float64};
                         seed::Union{Int,Nothing}=nothing, verbose::Bool=true, chain_offset::Int=0)
    n_sk <= 0 && throw(ArgumentError("n_sk must be positive. Current value: $n_sk"))
    d, C = size(xinit)
    (d == s.dim && size(vinit) == (d, C)) || throw(DimensionMismatch("xinit and vinit must have the same dimension as pdmp.dim ($(s.dim))"))
    X = Array{Float64}(undef, d, n_sk, C); V = similar(X)
    t = Array{Float64}(undef, n_sk, C); hz = similar(t); ar = similar(t)
    eva = Array{Float64}(undef, 5, n_sk, C)
    eb = Array{Int32}(undef, n_sk, C); rej = similar(eb); hh = similar(eb)
    status = zeros(Int32, C)
    sd = seed === nothing ? rand(UInt64) : UInt64(seed)
    GC.@preserve X V t hz ar eva eb rej hh status xinit vinit begin
        h = CHistory(pointer(X), pointer(V), pointer(t), pointer(hz), pointer(ar), pointer(eva), pointer(eb),
                     pointer(rej), pointer(hh), pointer(status), C_NULL, C_NULL, n_sk, 0)
        check(ccall((:pdmpflux_sample_skeleton, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, UInt64, Int64, Ptr{Cvoid}, Ref{CHistory}, Ptr{Cvoid}),
                    s.handle, C, n_sk, xinit, vinit, sd, chain_offset, C_NULL, Ref(h), C_NULL))
    end
    s.state = status
    return [PDMPHistory{Float64}(X[:, :, c], V[:, :, c], t[:, c], trues(d, n_sk), hz[:, c], ar[:, c], eb[:, c],
                                 eva[:, :, c], rej[:, c], hh[:, c]) for c in 1:C]
end
sample_skeleton(s::CuPDMP, n_sk::Int, xinit::Vector{Float64}, vinit::Vector{Float64}; kw...) =
    sample_skeleton(s, n_sk, reshape(xinit, :, 1), reshape(vinit, :, 1); kw...)[1]

"""
    sample_skeleton(sampler::CuPDMP, T::Float64, xinit, vinit; seed, init_capacity=1024) -> PDMPHistory

Replaces the time-horizon method (src/sample.jl:323-439): the skeleton ends with the point at exactly `t == T`.  The
number of events is not known in advance: like `_grow_history` (src/Composites.jl:172-191) the capacity doubles until
the chain fits (`PDMPFLUX_ERR_CAPACITY = -6`; the run is deterministic, so it is simply repeated).
"""
function sample_skeleton(s::CuPDMP, T::Float64, xinit::Vector{Float64}, vinit::Vector{Float64};
                         seed::Union{Int,Nothing}=nothing, verbose::Bool=true, init_capacity::Int=1024)
    (isfinite(T) && T >= 0) || throw(ArgumentError("T must be finite and non-negative. Current value: $T"))
    d = length(xinit)
    (d == s.dim && length(vinit) == d) || throw(DimensionMismatch("xinit and vinit must have the same dimension as pdmp.dim ($(s.dim))"))
    sd = seed === nothing ? rand(UInt64) : UInt64(seed)
    cap = max(1, init_capacity)
    while true
        X = Matrix{Float64}(undef, d, cap); V = similar(X)
        t = Vector{Float64}(undef, cap); hz = similar(t); ar = similar(t)
        eva = Matrix{Float64}(undef, 5, cap)
        eb = Vector{Int32}(undef, cap); rej = similar(eb); hh = similar(eb)
        status = zeros(Int32, 1); ncols = zeros(Int64, 1)
        rc = GC.@preserve X V t hz ar eva eb rej hh status ncols xinit vinit begin
            h = CHistory(pointer(X), pointer(V), pointer(t), pointer(hz), pointer(ar), pointer(eva), pointer(eb),
                         pointer(rej), pointer(hh), pointer(status), C_NULL, C_NULL, cap, 0)
            ccall((:pdmpflux_sample_skeleton_until, LIB), Cint,
                  (Ptr{Cvoid}, Int64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, UInt64, Int64, Ptr{Cvoid}, Ref{CHistory},
                   Ptr{Int64}, Ptr{Cvoid}),
                  s.handle, 1, T, cap, xinit, vinit, sd, 0, C_NULL, Ref(h), ncols, C_NULL)
        end
        if rc == -6      # PDMPFLUX_ERR_CAPACITY
            cap *= 2
            continue
        end
        check(rc)
        s.state = status
        n = Int(ncols[1])
        return PDMPHistory{Float64}(X[:, 1:n], V[:, 1:n], t[1:n], trues(d, n), hz[1:n], ar[1:n], eb[1:n], eva[:, 1:n],
                                    rej[1:n], hh[1:n])
    end
end

"Replaces `PDMPFlux.sample_from_skeleton(sampler, dt::Float64, history)` (src/sample.jl:573-646): samples at j*dt."
function sample_from_skeleton(s::CuPDMP, dt::Float64, h::PDMPHistory; discard_vt::Bool=true)
    (dt > 0 && isfinite(dt)) || throw(ArgumentError("dt must be positive. Current value: $dt"))
    d, n_sk = size(h.X)
    M = floor(Int, h.t[end] / dt)
    out = Matrix{Float64}(undef, discard_vt ? d : 2d + 1, M)
    M == 0 && return out
    check(ccall((:pdmpflux_sample_from_skeleton_dt, LIB), Cint,
                (Cint, Cint, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Int64, Int32,
                 Ptr{Float64}, Int32, Ptr{Cvoid}),
                s.flow_kind, d, n_sk, n_sk, 1, h.X, h.V, h.t, dt, M, discard_vt, out, 0, C_NULL))
    return out
end

"Replaces `PDMPFlux.sample_from_skeleton` (src/sample.jl:475-513)."
function sample_from_skeleton(s::CuPDMP, N::Int, h::PDMPHistory; discard_vt::Bool=true)
    N <= 0 && throw(ArgumentError("N must be positive. Current value: $N"))
    d, n_sk = size(h.X)
    out = Matrix{Float64}(undef, discard_vt ? d : 2d + 1, N)
    check(ccall((:pdmpflux_sample_from_skeleton, LIB), Cint,
                (Cint, Cint, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Ptr{Float64}, Int32, Ptr{Cvoid}),
                s.flow_kind, d, n_sk, 1, h.X, h.V, h.t, N, discard_vt, out, 0, C_NULL))
    return out
end

"""
Replaces `PDMPFlux.RV_diagnostic(history, U; B)` (src/diagnostic.jl:37-75): `U` is the value plugin of the sampler's
device potential.  `online=true` interpolates with the sampler's flow, i.e. the value
`sample_skeleton_with_diagnostic` (src/sample.jl:75-236) accumulates.
"""
function RV_diagnostic(h::PDMPHistory, s::CuPDMP; B::Int64=0, online::Bool=false)
    B < 0 && throw(ArgumentError("B must be non-negative. Current value: $B"))
    n_sk = length(h.t)
    n_sk == 0 && return 0.0
    T = h.t[end]
    (!isfinite(T) || T < 0.0) && throw(ArgumentError("history.t[end] must be finite and non-negative. Current value: $T"))
    rv = Ref{Float64}(0.0)
    check(ccall((:pdmpflux_rv_diagnostic, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ref{Float64}, Int32, Ptr{Cvoid}),
                s.pot, online ? s.flow_kind : Int32(0), n_sk, n_sk, 1, C_NULL, B, h.X, h.V, h.t, rv, 0, C_NULL))
    return rv[]
end

sample(s::CuPDMP, N_sk::Int, N::Int, xinit::Vector{Float64}, vinit::Vector{Float64}; seed=nothing, discard_vt=true) =
    sample_from_skeleton(s, N, sample_skeleton(s, N_sk, xinit, vinit; seed=seed); discard_vt=discard_vt)

end # module
