# record_tape.jl -- record the typed draw tape and the PDMPHistory of the REAL PDMPFlux.jl for a parity run.
#
# Julia is not installed in this repository's build image, so parity is established against the CPU oracle
# (oracle/).  Anyone with Julia + PDMPFlux.jl can pin the oracle (and the GPU path) against the reference itself:
#
#   julia tools/record_tape.jl out.bin          # writes draws (E, U, N) and the history in a flat binary
#   python tools/replay_tape.py out.bin         # (to be added by whoever has the file) replays it through
#                                               # oracle_c.sample_skeleton / pdmpflux_b200.sample_skeleton(tape=...)
#
# The only place the reference assigns `sampler.rng` is init_state (src/Samplers/AbstractPDMP.jl:100-101), so after
# init_state we swap in a recording AbstractRNG and drive get_event_state! directly (src/sample.jl:275-280).
using PDMPFlux, Random

mutable struct RecordingRNG <: Random.AbstractRNG
    inner::MersenneTwister
    E::Vector{Float64}; U::Vector{Float64}; N::Vector{Float64}
end
RecordingRNG(seed) = RecordingRNG(MersenneTwister(seed), Float64[], Float64[], Float64[])
Random.rand(r::RecordingRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) = (u = rand(r.inner); push!(r.U, u); u)
Random.randexp(r::RecordingRNG) = (e = randexp(r.inner); push!(r.E, e); e)
Random.randn(r::RecordingRNG) = (z = randn(r.inner); push!(r.N, z); z)
Random.randn(r::RecordingRNG, dims::Integer...) = (a = Array{Float64}(undef, dims...); for i in eachindex(a); a[i] = randn(r); end; a)
Random.randn!(r::RecordingRNG, a::AbstractArray{Float64}) = (for i in eachindex(a); a[i] = randn(r); end; a)

function record(path; dim=10, n_sk=10_001, seed=2024)
    U_Gauss(x) = sum(x .^ 2) / 2
    sampler = ZigZagAD(dim, U_Gauss)                       # README.md:40 (config C1)
    xinit, vinit = zeros(dim), ones(dim)
    state = PDMPFlux.init_state(sampler, xinit, vinit, seed)
    rng = RecordingRNG(seed)
    sampler.rng = rng
    history = PDMPHistory(dim, n_sk)
    PDMPFlux.record!(history, 1, state, dim)
    for k in 2:n_sk
        state = PDMPFlux.get_event_state(state, sampler)
        PDMPFlux.record!(history, k, state, dim)
    end
    open(path, "w") do io
        write(io, Int64[dim, n_sk, length(rng.E), length(rng.U), length(rng.N)])
        write(io, rng.E); write(io, rng.U); write(io, rng.N)
        write(io, history.X); write(io, history.V); write(io, history.t); write(io, history.horizon); write(io, history.ar)
        write(io, history.errored_bound); write(io, history.rejected); write(io, history.hitting_horizon)
    end
end

record(length(ARGS) >= 1 ? ARGS[1] : "pdmpflux_tape.bin")
