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
"""One rank's shard: `packed_rows` holds only rows [row_begin, row_end) of every column (cld(rows, 4) bytes each);
the marker statistics are completed by init_sharding! (integer code counts summed over the ranks)."""
function GpuBackendShard(packed_rows::Matrix{UInt8}, nObs::Integer, ntraits::Integer, row_begin::Integer, row_end::Integer;
                         device::Integer=0)
    stride, p = size(packed_rows)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve packed_rows check(ccall((:jwas_create_shard, LIB), Cint,
        (Int64, Int64, Cint, Int64, Int64, Ptr{UInt8}, Int64, Cint, Ref{Ptr{Cvoid}}),
        nObs, p, ntraits, row_begin, row_end, packed_rows, stride, device, h))
    b = GpuBackend(h[], nObs, p, ntraits, Vector{Float32}(undef, p), Vector{Float32}(undef, p), nothing)
    for (key, val) in (("engine", 1), ("lag", 2), ("chain_ctas", 6))
        check(ccall((:jwas_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), h[], key, val))
    end
    finalizer(x -> ccall((:jwas_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), b)
    return b        # marker_means / xpRinvx are filled by init_sharding!
end
init_sharding!(b::GpuBackend, rank::Integer, world::Integer, nccl_id::Vector{UInt8}) =
    (check(ccall((:jwas_init_sharding, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), b.handle, rank, world, nccl_id));
     check(ccall((:jwas_get_marker_stats, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}), b.handle, b.marker_means, b.xpRinvx)))

"BayesC with scalar σ²α and π (MCMC_BayesianAlphabet.jl:231 fills the vector on the host; here the fill is on the device)"
function BayesC_gpu!(b::GpuBackend, vare, varEffect, π; schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1)
    st = Ref{SweepStats}()
    check(ccall((:jwas_sweep_bayesc, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Cdouble, UInt64, UInt32, Ref{SweepStats}),
        b.handle, schedule, Float64(vare), Float64(varEffect), Float64(π), seed, UInt32(iter), st))
    b.last_stats = st[]
    return nothing
end

# marker-level priors (annotated runs: genotypes.annotations.snp_pi, nMarkers x K, BayesR.jl:28, MTBayesABC.jl:28-30)
# go to the C side row-major; a K-vector is the usual global prior
prior_arg(π::Vector{Float64}) = (π, Cint(0))
prior_arg(snp_pi::Matrix{Float64}) = (collect(permutedims(snp_pi)), Cint(1))

"""Drop-in for BayesR!(genotypes, ycorr, vare) / BayesR_block! (BayesR.jl:27-43); `π` is the 4-vector, or the
nMarkers x 4 matrix `genotypes.annotations.snp_pi` of an annotated run (BayesR.jl:28)."""
function BayesR_gpu!(b::GpuBackend, vare, sigmaSq, π::Union{Vector{Float64},Matrix{Float64}}, gamma::Vector{Float64};
                     schedule=SCHED_EXACT, iter::Integer=1, burnin::Integer=0, seed::UInt64=UInt64(0))
    st = Ref{SweepStats}()
    full = iter <= burnin ? Cint(0) : Cint(1)                     # bayesr_block_nreps, BayesR.jl:22-25
    pr, per_marker = prior_arg(π)
    GC.@preserve pr gamma check(ccall((:jwas_sweep_bayesr, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble, Ptr{Float64}, Cint, Ptr{Float64}, Cint, UInt64, UInt32,
         Ptr{Float64}, Ptr{Float64}, Ref{SweepStats}),
        b.handle, schedule, full, Float64(vare), Float64(sigmaSq), pr, per_marker, gamma, Cint(length(gamma)), seed,
        UInt32(iter), C_NULL, C_NULL, st))
    b.last_stats = st[]
    return nothing
end

"""Drop-in for MTBayesABC!(genotypes, wArray, vare, locus_effect_variances, nModels) with sampler I (MTBayesABC.jl:37-54).
`bigPi`: the 2^t joint-state priors indexed sum(δ_k << (k-1)) -- for two traits 00, 10, 01, 11, the column order of
`annotations.snp_pi` (annotation_setup.jl:15), whose nMarkers x 4 matrix is accepted as it is."""
function MTBayesABC_gpu!(b::GpuBackend, R::Matrix{Float64}, G::Matrix{Float64}, bigPi::Union{Vector{Float64},Matrix{Float64}};
                         schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1)
    st = Ref{SweepStats}()
    Rr = collect(permutedims(R)); Gr = collect(permutedims(G))     # row-major for the C side
    pr, per_marker = prior_arg(bigPi)
    GC.@preserve Rr Gr pr check(ccall((:jwas_sweep_mt1, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Cint, UInt64, UInt32, Ptr{Float64}, Ptr{Float64}, Ref{SweepStats}),
        b.handle, schedule, Rr, Gr, Cint(0), pr, per_marker, seed, UInt32(iter), C_NULL, C_NULL, st))
    b.last_stats = st[]
    return nothing
end

"Sampler II, two traits (MTBayesABC.jl:129-210, 439-646); bigPi in the order 00, 10, 01, 11"
function MTBayesABC_II_gpu!(b::GpuBackend, R::Matrix{Float64}, G::Matrix{Float64}, bigPi::Vector{Float64};
                            schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1)
    st = Ref{SweepStats}()
    Rr = collect(permutedims(R)); Gr = collect(permutedims(G))
    GC.@preserve Rr Gr bigPi check(ccall((:jwas_sweep_mt2, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt64, UInt32, Ptr{Float64}, Ptr{Float64}, Ref{SweepStats}),
        b.handle, schedule, Rr, Gr, bigPi, seed, UInt32(iter), C_NULL, C_NULL, st))
    b.last_stats = st[]
    return nothing
end

"megaBayesABC! (BayesABC.jl:1-7; constraint=true): per-trait vare, σ²α, π -- the traits share one column read"
function megaBayesABC_gpu!(b::GpuBackend, vare::Vector{Float64}, varEffects::Vector{Float64}, π::Vector{Float64};
                           schedule=SCHED_EXACT, seed::UInt64=UInt64(0), iter::Integer=1)
    st = Ref{SweepStats}()
    GC.@preserve vare varEffects π check(ccall((:jwas_sweep_mega, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, UInt64, UInt32, Ptr{Float64}, Ptr{Float64}, Ref{SweepStats}),
        b.handle, schedule, vare, varEffects, π, seed, UInt32(iter), C_NULL, C_NULL, st))
    b.last_stats = st[]
    return nothing
end

"BayesB: G.val stays a device vector (MCMC_BayesianAlphabet.jl:67-69); its χ² update runs there (variance_components.jl:169-172)"
fill_var_effects!(b::GpuBackend, v) = check(ccall((:jwas_fill_hyper, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble), b.handle, Cint(0), Float64(v)))
fill_pi!(b::GpuBackend, π) = check(ccall((:jwas_fill_hyper, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble), b.handle, Cint(1), Float64(π)))
sample_bayesb_variances!(b::GpuBackend, df, scale; seed::UInt64=UInt64(0), iter::Integer=1) =
    check(ccall((:jwas_sample_bayesb_variances, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, UInt64, UInt32, Ptr{Float64}),
                b.handle, Float64(df), Float64(scale), seed, UInt32(iter), C_NULL))

"ycorr -= M*α for the α on the device (MCMC_BayesianAlphabet.jl:137-143); EBV = M*α of one trait (output.jl:302)"
ycorr_sub_malpha!(b::GpuBackend) = check(ccall((:jwas_ycorr_sub_malpha, LIB), Cint, (Ptr{Cvoid},), b.handle))
function mul_alpha(b::GpuBackend, trait::Integer)
    out = Vector{Float32}(undef, b.nObs)
    check(ccall((:jwas_mul_alpha, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}), b.handle, Cint(trait - 1), out))
    return out
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

function get_means(b::GpuBackend)
    tp = b.ntraits * b.nMarkers
    m = Vector{Float32}(undef, tp); m2 = Vector{Float32}(undef, tp); md = Vector{Float32}(undef, tp)
    check(ccall((:jwas_get_means, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), b.handle, m, m2, md))
    return m, m2, md
end

# ---- several GPUs, one Julia process per GPU (Distributed / MPI.jl carry the two small byte strings) ----
"rank 0 makes the 128-byte NCCL id; broadcast it, then every rank calls init_sharding!"
function nccl_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:jwas_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
"after set_blocks!: export this rank's 64-byte exchange-buffer handle, all-gather them in rank order, import"
function ipc_export(b::GpuBackend)
    hd = Vector{UInt8}(undef, 64)
    check(ccall((:jwas_ipc_export, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), b.handle, hd))
    return hd
end
ipc_import!(b::GpuBackend, handles::Vector{UInt8}) =
    check(ccall((:jwas_ipc_import, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), b.handle, handles))

end # module
