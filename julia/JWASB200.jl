# JWASB200.jl -- reference-side binding of libjwasb200.so (include/jwas_b200.h).
#
# SOURCE ONLY: Julia is not installed in the build image or on the GPU box, so this file has not
# been executed there.  It is the `ccall` layer a JWAS.jl maintainer adds next to
# src/1.JWAS/src/markers/streaming_genotypes.jl to get a third storage mode (`storage=:gpu`) that
# goes through the same seam as `storage=:stream` (types.jl:149-150; readgenotypes.jl:236-295;
# MCMC/MCMC_BayesianAlphabet.jl:53-65, 243-251).  Every wrapper mutates the same Genotypes fields
# the CPU samplers mutate and returns `nothing`, like BayesABC!/BayesR!/MTBayesABC!.
module JWASB200

const LIB = get(ENV, "JWAS_B200_LIB", "libjwasb200.so")

const SCHED_EXACT, SCHED_BLOCK, SCHED_INDEPENDENT = Cint(0), Cint(1), Cint(2)

# mirrors jwas_sweep_stats (include/jwas_b200.h), field for field
struct SweepStats
    ycorr_ss::NTuple{16,Cdouble}
    ycorr_sum::NTuple{4,Cdouble}
    alpha_ss::NTuple{16,Cdouble}
    beta_ss::NTuple{16,Cdouble}
    nnz_alpha::NTuple{4,Cdouble}
    sum_delta::NTuple{4,Cdouble}
    class_counts::NTuple{16,Cdouble}
    bayesr_ssq::Cdouble
    ycorr_maxabs::Cdouble
    scale_exp::Int32
    overflow::Int32
    n_active::Int64
    n_rounds::Int64
end

mutable struct GpuBackend               # what Genotypes.stream_backend holds for storage=:gpu
    handle::Ptr{Cvoid}
    nObs::Int
    nMarkers::Int
    ntraits::Int
    marker_means::Vector{Float32}
    xpRinvx::Vector{Float32}
    last_stats::Union{Nothing,SweepStats}
end

check(rc::Cint) = rc == 0 ? nothing :
    error(unsafe_string(ccall((:jwas_last_error, LIB), Cstring, ())))   # ErrorException, like error("...")

"""Replaces GibbsMats(...) / load_streaming_backend: `packed` is the .jgb2 image
(p columns of cld(nObs,4) bytes, streaming_genotypes.jl:622-627)."""
function GpuBackend(packed::Matrix{UInt8}, nObs::Integer, ntraits::Integer; device::Integer=0)
    stride, p = size(packed)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve packed check(ccall((:jwas_create, LIB), Cint,
        (Int64, Int64, Cint, Ptr{UInt8}, Int64, Cint, Ref{Ptr{Cvoid}}),
        nObs, p, ntraits, packed, stride, device, h))
    means = Vector{Float32}(undef, p); xpx = Vector{Float32}(undef, p)
    check(ccall((:jwas_get_marker_stats, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}), h[], means, xpx))
    b = GpuBackend(h[], nObs, p, ntraits, means, xpx, nothing)
    # backend tuning (no effect on results): persistent fused kernel, lag-2 exact schedule, chain pipelined over
    # six chain CTAs -- the configuration bench.py measures (panels of 4096 markers).  Must precede set_blocks!.
    for (key, val) in (("engine", 1), ("lag", 2), ("chain_ctas", 6))
        check(ccall((:jwas_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), h[], key, val))
    end
    finalizer(x -> ccall((:jwas_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), b)   # like streaming_genotypes.jl:966-968
    return b
end

"fast_blocks block starts (1-based, JWAS.jl:293-316) -> Gram blocks on the device"
function set_blocks!(b::GpuBackend, block_starts::Vector{Int})
    bounds = Int64[block_starts .- 1; b.nMarkers]
    check(ccall((:jwas_set_blocks, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), b.handle, bounds, length(block_starts)))
end

put_ycorr!(b::GpuBackend, ycorr::Vector{Float32}) =
    check(ccall((:jwas_put_ycorr, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}), b.handle, ycorr))
get_ycorr!(ycorr::Vector{Float32}, b::GpuBackend) =
    check(ccall((:jwas_get_ycorr, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}), b.handle, ycorr))
put_state!(b::GpuBackend, α::Vector{Float32}, β::Vector{Float32}, δ::Vector{Int32}) =
    check(ccall((:jwas_put_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}), b.handle, α, β, δ))
get_state!(α::Vector{Float32}, β::Vector{Float32}, δ::Vector{Int32}, b::GpuBackend) =
    check(ccall((:jwas_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}), b.handle, α, β, δ))

"""Drop-in for BayesABC!(genotypes, ycorr, vare, locus_effect_variances) (BayesABC.jl:10-15) with
device-resident ycorr/α/β/δ: nothing crosses PCIe but a few scalars.  `u`,`z` may carry draws generated
by the host in reference order (rand(), randn() per marker, BayesABC.jl:44,46,54); `nothing` uses the
library's Philox stream."""
function BayesABC_gpu!(b::GpuBackend, vare, varEffects::Vector{Float64}, π::Vector{Float64};
                       schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1, u=nothing, z=nothing)
    st = Ref{SweepStats}()
    pu = u === nothing ? Ptr{Float64}(C_NULL) : pointer(u)
    pz = z === nothing ? Ptr{Float64}(C_NULL) : pointer(z)
    GC.@preserve varEffects π u z check(ccall((:jwas_sweep_bayesabc, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cdouble, Ptr{Float64}, Ptr{Float64}, UInt64, UInt32, Ptr{Float64}, Ptr{Float64}, Ref{SweepStats}),
        b.handle, schedule, Float64(vare), varEffects, π, seed, UInt32(iter), pu, pz, st))
    b.last_stats = st[]
    return nothing
end

"""The reference call itself on HOST arrays: BayesABC!(xArray, xRinvArray, xpRinvx, yCorr, α, β, δ, vare, varEffects, π)
mutates yCorr, α, β, δ in place (BayesABC.jl:60-63).  One ccall: copies in, the sweep, copies out."""
function BayesC_host!(b::GpuBackend, yCorr::Vector{Float32}, α::Vector{Float32}, β::Vector{Float32}, δ::Vector{Int32},
                      vare, varEffect, π; schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1)
    st = Ref{SweepStats}()
    GC.@preserve yCorr α β δ check(ccall((:jwas_sweep_bayesc_host, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Cdouble, UInt64, UInt32, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32},
         Ref{SweepStats}),
        b.handle, schedule, Float64(vare), Float64(varEffect), Float64(π), seed, UInt32(iter), yCorr, α, β, δ, st))
    b.last_stats = st[]
    return nothing
end

"""Centre on the means get_genotypes computed on ALL genotyped individuals (readgenotypes.jl:372-385) when the rows
uploaded are the phenotyped subset (JWAS.jl:381-402).  Before set_blocks!."""
set_marker_means!(b::GpuBackend, means::Vector{Float32}) =
    (check(ccall((:jwas_set_marker_means, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}), b.handle, means));
     check(ccall((:jwas_get_marker_stats, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}), b.handle, b.marker_means, b.xpRinvx)))

"""Rows [first, last) (0-based) of the packed columns that rank `rank` of `world` GPU processes stores
(one process per GPU; Distributed / MPI.jl move the 128-byte NCCL id and the 64-byte IPC handles)."""
function shard_range(nObs::Integer, rank::Integer, world::Integer)
    b = Ref{Int64}(0); e = Ref{Int64}(0)
    check(ccall((:jwas_shard_range, LIB), Cint, (Int64, Cint, Cint, Ref{Int64}, Ref{Int64}), nObs, rank, world, b, e))
    return b[], e[]
end
init_sharding!(b::GpuBackend, rank::Integer, world::Integer, nccl_id::Vector{UInt8}) =
    check(ccall((:jwas_init_sharding, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), b.handle, rank, world, nccl_id))

"BayesC with scalar σ²α and π (MCMC_BayesianAlphabet.jl:231 fills the vector on the host; here the fill is on the device)"
function BayesC_gpu!(b::GpuBackend, vare, varEffect, π; schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1)
    st = Ref{SweepStats}()
    check(ccall((:jwas_sweep_bayesc, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Cdouble, UInt64, UInt32, Ref{SweepStats}),
        b.handle, schedule, Float64(vare), Float64(varEffect), Float64(π), seed, UInt32(iter), st))
    b.last_stats = st[]
    return nothing
end

"Drop-in for BayesR!(genotypes, ycorr, vare) / BayesR_block! (BayesR.jl:27-43)"
function BayesR_gpu!(b::GpuBackend, vare, sigmaSq, π::Vector{Float64}, gamma::Vector{Float64};
                     schedule=SCHED_EXACT, iter::Integer=1, burnin::Integer=0, seed::UInt64=UInt64(0))
    st = Ref{SweepStats}()
    full = iter <= burnin ? Cint(0) : Cint(1)                     # bayesr_block_nreps, BayesR.jl:22-25
    check(ccall((:jwas_sweep_bayesr, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Cint, UInt64, UInt32,
         Ptr{Float64}, Ptr{Float64}, Ref{SweepStats}),
        b.handle, schedule, full, Float64(vare), Float64(sigmaSq), π, Cint(0), gamma, Cint(length(gamma)), seed,
        UInt32(iter), C_NULL, C_NULL, st))
    b.last_stats = st[]
    return nothing
end

"Drop-in for MTBayesABC!(genotypes, wArray, vare, locus_effect_variances, nModels) with sampler I (MTBayesABC.jl:37-54)"
function MTBayesABC_gpu!(b::GpuBackend, R::Matrix{Float64}, G::Matrix{Float64}, bigPi::Vector{Float64};
                         schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1)
    st = Ref{SweepStats}()
    Rr = collect(permutedims(R)); Gr = collect(permutedims(G))     # row-major for the C side
    check(ccall((:jwas_sweep_mt1, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, UInt64, UInt32, Ptr{Float64}, Ptr{Float64}, Ref{SweepStats}),
        b.handle, schedule, Rr, Gr, Cint(0), bigPi, Cint(0), seed, UInt32(iter), C_NULL, C_NULL, st))
    b.last_stats = st[]
    return nothing
end

"intercept-only location update without moving ycorr (MCMC_BayesianAlphabet.jl:207-220)"
function shift_ycorr!(b::GpuBackend, trait::Integer, shift::Real)
    s = Ref{Cdouble}(0); ss = Ref{Cdouble}(0)
    check(ccall((:jwas_shift_ycorr, LIB), Cint, (Ptr{Cvoid}, Cint, Cfloat, Ref{Cdouble}, Ref{Cdouble}),
        b.handle, Cint(trait - 1), Float32(shift), s, ss))
    return s[], ss[]
end

"output_posterior_mean_variance for α, α², δ (output.jl:568-577)"
accumulate!(b::GpuBackend, nsamples; bayesr::Bool=false) =
    check(ccall((:jwas_accumulate, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cint), b.handle, Float64(nsamples), Cint(bayesr)))

end # module
