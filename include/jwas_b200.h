/*
 * jwas_b200.h -- C ABI of libjwasb200.so, the B200 (sm_100a) backend for the JWAS
 * marker-effects Gibbs sweep.
 *
 * JWAS.jl has no plugin interface; the seam it used for its own second backend is
 * Genotypes.storage_mode / Genotypes.stream_backend plus a per-backend sampler picked
 * in the MCMC loop (types.jl:149-150; MCMC/MCMC_BayesianAlphabet.jl:53-65, 243-251;
 * markers/readgenotypes.jl:236-295).  These entry points are what a third
 * `storage=:gpu` branch binds with `ccall`; each one cites the
 * reference routine it replaces.  Paths are relative to src/1.JWAS/src/.
 *
 * Conventions
 *   - plain C, no C++/torch types; every call returns 0 on success, non-zero on error,
 *     and jwas_last_error() describes the last failure of the calling thread
 *     (the reference raises ErrorException via error("..."); wrappers re-raise).
 *   - the library owns device memory behind the handle; host buffers are borrowed only
 *     for the duration of the (synchronous) call.
 *   - one host thread per handle; no callbacks into the host language.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *   - markers are 0-based here; block starts are 0-based boundaries (nblocks+1 entries).
 *   - unit residual weights only (Rinv == 1): heterogeneous_residuals is rejected by the
 *     host wrapper before it gets here.
 */
#ifndef JWAS_B200_H
#define JWAS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct jwas_handle jwas_handle;

/* sweep schedule -- which reference sampler the call reproduces */
#define JWAS_SCHED_EXACT        0  /* BayesABC! / BayesR! / _MTBayesABC_samplerI!: one pass, marker by marker.
                                      Runs as look-ahead panels with one repetition (the in-block Gram
                                      identity of BayesABC_block!, BayesABC.jl:157,169). */
#define JWAS_SCHED_BLOCK        1  /* BayesABC_block! etc. with independent_blocks=false: nreps = block size */
#define JWAS_SCHED_INDEPENDENT  2  /* BayesABC_block_independent! etc.: all blocks read a ycorr snapshot */

/* reductions a sweep returns so the host-side hyper-parameter draws need no extra pass
 * (variance_components.jl:60-79, 82-112, 151-189; Pi.jl:7-42; MCMC_BayesianAlphabet.jl:357-365) */
typedef struct {
    double ycorr_ss[16];     /* t*t row-major: ycorr_i' ycorr_j                       */
    double ycorr_sum[4];     /* per trait sum(ycorr)                                  */
    double alpha_ss[16];     /* t*t: alpha_i' alpha_j (BayesC single trait: alpha'alpha) */
    double beta_ss[16];      /* t*t: beta_i' beta_j (multi-trait BayesC variance update) */
    double nnz_alpha[4];     /* per trait #(alpha != 0)                               */
    double sum_delta[4];     /* per trait sum(delta) (ABC); BayesR: #(delta > 1)      */
    double class_counts[16]; /* BayesR: counts per class; MT: counts per joint state  */
    double bayesr_ssq;       /* sum alpha_j^2 / gamma[delta_j] over delta_j > 1       */
    double ycorr_maxabs;     /* max |ycorr| (next sweep's fixed-point scale)          */
    int32_t scale_exp;       /* S used by this sweep's fixed-point dots               */
    int32_t overflow;        /* 1 if the fixed-point clamp was hit (call also fails)  */
    int64_t n_active;        /* marker updates with delta-alpha != 0 (sequential depth) */
    int64_t n_rounds;        /* speculative rounds executed by the chain              */
} jwas_sweep_stats;

/* ---- lifetime: replaces GibbsMats(...) (markers/tools4genotypes.jl:237-275) and
 *      load_streaming_backend (markers/streaming_genotypes.jl:884-971) -------------- */
/* packed: marker-major 2-bit codes exactly as in a .jgb2 file (p columns of `stride`
 * bytes; individual i in byte i>>2, bits (i&3)<<1; code 3 = missing).  Column means and
 * xpRinvx are computed on the device. */
int jwas_create(int64_t n_obs, int64_t n_markers, int n_traits,
                const uint8_t* packed, int64_t stride_bytes, int device, jwas_handle** out);
/* Synthetic genotypes generated directly as packed bytes on the device (never materialised
 * dense): f_j ~ U(0.05,0.5), code = Bernoulli(f_j)+Bernoulli(f_j), optional missing rate;
 * Philox keyed by `seed` (benchmarks/bayesr_parity_common.jl:28-59 is the model). */
int jwas_create_synthetic(int64_t n_obs, int64_t n_markers, int n_traits, uint64_t seed,
                          double missing_rate, int device, jwas_handle** out);
/* Row shards (multi-GPU, one process per GPU): rank r of `world` stores rows [begin,end) of every column --
 * words of 64 individuals split evenly (jwas_shard_range).  A shard is created from the bytes of its own rows
 * only (column pitch stride_bytes >= cld(end-begin,4)); its marker statistics are completed by
 * jwas_init_sharding, which sums the integer code counts over the ranks. */
int jwas_shard_range(int64_t n_obs, int rank, int world, int64_t* begin, int64_t* end);
int jwas_create_shard(int64_t n_obs, int64_t n_markers, int n_traits, int64_t row_begin, int64_t row_end,
                      const uint8_t* packed_rows, int64_t stride_bytes, int device, jwas_handle** out);
int jwas_create_synthetic_shard(int64_t n_obs, int64_t n_markers, int n_traits, uint64_t seed, double missing_rate,
                                int64_t row_begin, int64_t row_end, int device, jwas_handle** out);
/* copy of the packed rows stored by this handle back to the host (p * stride bytes, stride = cld(rows,4)) */
int jwas_get_packed(jwas_handle* h, uint8_t* packed, int64_t stride_bytes);
/* Centre on means computed elsewhere: get_genotypes centres on ALL genotyped individuals
 * (markers/readgenotypes.jl:372-385) before the rows are aligned to the phenotyped ones (JWAS.jl:381-402).
 * xpRinvx is recomputed for these means.  Call before jwas_set_blocks. */
int jwas_set_marker_means(jwas_handle* h, const float* means);
int jwas_destroy(jwas_handle* h);
const char* jwas_last_error(void);
int jwas_device_count(void);

/* marker_means / xpRinvx (Packed2BitBackend fields, streaming_genotypes.jl:16-17) */
int jwas_get_marker_stats(jwas_handle* h, float* means, float* xpx);

/* block partition + Gram blocks X_b'X_b: GibbsMats(...; fast_blocks) (tools4genotypes.jl:259-269).
 * Must be called before any sweep (JWAS_SCHED_EXACT uses it as the look-ahead panel partition). */
int jwas_set_blocks(jwas_handle* h, const int64_t* starts, int64_t nblocks);

/* ---- ycorr lifecycle (MCMC/MCMC_BayesianAlphabet.jl:131-147, 207-220, 365) ---------- */
int jwas_put_ycorr(jwas_handle* h, const float* ycorr /* t*n */);
int jwas_get_ycorr(jwas_handle* h, float* ycorr);
/* ycorr[trait] -= M * alpha[trait] for the current device alpha (:137-143, streaming_mul_alpha!) */
int jwas_ycorr_sub_malpha(jwas_handle* h);
/* ycorr[trait] += shift (intercept-only location update :207-220 without moving ycorr);
 * returns sum and sum of squares after the shift */
int jwas_shift_ycorr(jwas_handle* h, int trait, float shift, double* sum, double* sumsq);
/* out = M * alpha[trait] (getEBV, output.jl:281-306) */
int jwas_mul_alpha(jwas_handle* h, int trait, float* out /* n */);

/* ---- sampler state alpha, beta, delta (Genotypes fields, types.jl:124-131) ----------- */
int jwas_put_state(jwas_handle* h, const float* alpha, const float* beta, const int32_t* delta);
int jwas_get_state(jwas_handle* h, float* alpha, float* beta, int32_t* delta);

/* ---- sweeps.  u/z: optional replayed draw tables indexed [(rep*t + trait)*p + marker]
 *      (host generates them in reference order); NULL -> native Philox stream
 *      keyed (seed; marker, iter, slot, rep) of include/jwas_contract.h ---------------- */
/* BayesABC! (BayesABC.jl:60-80), BayesABC_block! (:118-188), _independent! (:190-255).
 * var_effects: p entries (BayesC passes fill(G.val,p), MCMC_BayesianAlphabet.jl:231);
 * pi: p entries = P(effect is zero) (BayesABC.jl:66-68). */
int jwas_sweep_bayesabc(jwas_handle* h, int schedule, double vare,
                        const double* var_effects, const double* pi,
                        uint64_t seed, uint32_t iter, const double* u, const double* z,
                        jwas_sweep_stats* stats);
/* scalar convenience forms (device-side fill) */
int jwas_sweep_bayesc(jwas_handle* h, int schedule, double vare, double var_effect, double pi,
                      uint64_t seed, uint32_t iter, jwas_sweep_stats* stats);
/* The reference call itself -- BayesABC!(xArray, xRinvArray, xpRinvx, yCorr, alpha, beta, delta, vare, varEffects, pi)
 * mutating the CALLER's host arrays in place (BayesABC.jl:60-63): host -> device copies, the sweep, device -> host
 * copies, enqueued back to back with one synchronisation (use pinned buffers for asynchronous copies). */
int jwas_sweep_bayesc_host(jwas_handle* h, int schedule, double vare, double var_effect, double pi,
                           uint64_t seed, uint32_t iter, float* ycorr, float* alpha, float* beta, int32_t* delta,
                           jwas_sweep_stats* stats);
/* BayesR! (BayesR.jl:45-97), BayesR_block! (:111-193), _independent! (:195-273).
 * full_reps: the burn-in gate of bayesr_block_nreps (:22-25) evaluated by the caller. */
int jwas_sweep_bayesr(jwas_handle* h, int schedule, int full_reps, double vare, double sigma_sq,
                      const double* pi, int per_marker_pi, const double* gamma, int nclasses,
                      uint64_t seed, uint32_t iter, const double* u, const double* z,
                      jwas_sweep_stats* stats);
/* _MTBayesABC_samplerI! (MTBayesABC.jl:57-127), block (:243-333), independent (:335-437).
 * R, G: t*t row-major covariances (G: p*t*t if per_marker_G);
 * big_pi: 2^t joint-state priors indexed sum(delta_k << k) (or p*2^t if per_marker_pi). */
int jwas_sweep_mt1(jwas_handle* h, int schedule, const double* R, const double* G, int per_marker_G,
                   const double* big_pi, int per_marker_pi,
                   uint64_t seed, uint32_t iter, const double* u, const double* z,
                   jwas_sweep_stats* stats);

/* device-resident hyper-parameter vectors used when a sweep is called with NULL pointers:
 * which = 0 -> var_effects[j] = value for all j (BayesB start, MCMC_BayesianAlphabet.jl:67-69);
 * which = 1 -> pi[j] = value (bayesabc_pi_vector, BayesABC.jl:17-23) */
int jwas_fill_hyper(jwas_handle* h, int which, double value);

/* _MTBayesABC_samplerII! (MTBayesABC.jl:129-210; block :439-537, independent :539-646): joint draw of the
 * 2^t inclusion states, t = 2.  big_pi: 4 joint-state priors in the order 00,10,01,11. */
int jwas_sweep_mt2(jwas_handle* h, int schedule, const double* R, const double* G, const double* big_pi,
                   uint64_t seed, uint32_t iter, const double* u, const double* z, jwas_sweep_stats* stats);

/* megaBayesABC! (BayesABC.jl:1-7; constraint=true, MCMC_BayesianAlphabet.jl:233-234): one single-trait BayesABC
 * step per trait with vare[k], var_effects[k], pi[k] (t entries each) -- the traits share one column read. */
int jwas_sweep_mega(jwas_handle* h, int schedule, const double* vare, const double* var_effects, const double* pi,
                    uint64_t seed, uint32_t iter, const double* u, const double* z, jwas_sweep_stats* stats);

/* BayesB per-marker variance update on device (variance_components.jl:169-172):
 * var_j = (beta_j^2 + df*scale) / chisq(df+1), chi-square from the native stream. */
int jwas_sample_bayesb_variances(jwas_handle* h, double df, double scale, uint64_t seed,
                                 uint32_t iter, double* var_effects_out /* p, may be NULL */);

/* ---- posterior accumulators (output.jl:556-577) --------------------------------------- */
int jwas_accumulate(jwas_handle* h, double nsamples, int bayesr /* meanDelta tracks delta>1 */);
int jwas_get_means(jwas_handle* h, float* mean_alpha, float* mean_alpha2, float* mean_delta);

/* Gram block of block `ib` (b*b floats, row-major) -- XpRinvX[ib] of GibbsMats */
int jwas_get_gram(jwas_handle* h, int64_t ib, float* out);

/* ---- multi-GPU: individuals (rows of M) shard across the GPUs of one node, one process per GPU.
 * Rank r STORES and streams only its rows of every column (jwas_shard_range); ycorr and the sampler state
 * (alpha, beta, delta, Gram blocks, marker statistics) are replicated.  Per marker block the exact int64
 * partial rhs of every rank are summed, every rank runs the identical chain (same draws -> same bits) and
 * applies the axpy to its own rows; the ycorr shards are all-gathered once at the end of the sweep (the host's
 * hyper-parameter draws need ycorr'ycorr).  Results are bit-identical for any number of ranks.
 * Order of calls: create (full matrix, or this rank's shard) -> jwas_init_sharding -> jwas_set_blocks ->
 * jwas_ipc_export / all-gather the 64-byte handles / jwas_ipc_import -> sweeps.
 * rank 0 calls jwas_nccl_unique_id and the host language broadcasts the 128 bytes (torch.distributed, MPI ...). */
int jwas_nccl_unique_id(uint8_t* out128);
/* A handle created from the full matrix keeps only its own rows from here on; a shard is checked against
 * jwas_shard_range.  Marker statistics are summed over the ranks (NCCL, integer counts: exact). */
int jwas_init_sharding(jwas_handle* h, int rank, int world, const uint8_t* unique_id128);
int jwas_get_row_range(jwas_handle* h, int64_t* begin, int64_t* end);
/* Fused multi-GPU sweep (engine 1, lag 1): every rank exports its exchange buffer as a CUDA IPC handle
 * (64 bytes), the host language all-gathers the handles, every rank imports them.  The persistent
 * kernel then pushes each block's exact int64 partial rhs straight into the peers' memory over NVLink as
 * self-validating 16-byte words {lo32, tag, hi32, tag} (one communication CTA per GPU; no flag, no fence,
 * no NCCL call, no kernel boundary inside the sweep); the chain's rhs read polls the tags.  Without the
 * import the sharded sweep falls back to engine 0 with one NCCL all-reduce per block. */
int jwas_ipc_export(jwas_handle* h, uint8_t* out64);
int jwas_ipc_import(jwas_handle* h, const uint8_t* handles /* world * 64 bytes in rank order */);

/* ---- introspection used by bench.py / tests ------------------------------------------- */
int64_t jwas_kernel_launches(jwas_handle* h);     /* kernels launched by this handle so far */
/* Backend options (none changes a result bit):
 *   "engine"      0 = multi-kernel engine, 1 = persistent fused sweep kernel
 *   "lag"         L = 1, 2: lagged exact schedule (engine 1): the stream of panel k carries the updates of panels <= k-L-1,
 *                 the chains of panels k-L..k-1 overlap it (cross-Gram corrections); 2 needs chain_ctas >= 1
 *   "chain_ctas"  engine 1, lag 1: number of chain CTAs of the PIPELINED chain (units of <= 1024 markers handed
 *                 from CTA to CTA as 64-bit commit records); 0 = one chain CTA.  Re-cuts the row slices.
 *   "gather"      pipelined chain: 1 = one warp of every streaming CTA replays the commit records under the
 *                 stream (pays off with panels that are a multiple of 31*16 markers), 0 = in line (default)
 *   "ws"          engine 1, pipelined chain, one trait without missing calls: 1 (default) = warp-specialised streaming
 *                 role (builder warps rebuild one lookup-table set while the streaming warps run through the other;
 *                 per-warp release of a panel), 0 = the plain role (record replay, rebuild, stream in turn)
 *   "l2_prefetch" 1 = pull the next panel's tile into L2 at the end of a panel (default 0: measured slower)
 *   "poll_ns_stream", "poll_ns_chain"  back-off in ns after an empty poll of a commit record (default 0)
 *   "stream_variant", "stream_pf"   independent schedule, engine 1: launch shape of the streamed block-rhs kernel
 *                 (0 = 512 threads + register double buffer, 1 = 1024 threads, 2 = 768 + double buffer,
 *                 3 = 1024 + double buffer) and its L2 prefetch distance in chunk iterations (0 = off)
 *   "profile"     1 = time the streaming kernel(s) with CUDA events (jwas_last_stream_kernel_ms)
 *   "timers"      1 = in-kernel phase timers; only in a library built with -DJW_TIMERS
 *   "gram_popcount" 1 = popcount Gram kernel instead of the bf16 tensor-core GEMM */
int jwas_set_option(jwas_handle* h, const char* key, int64_t value);
/* last sweep's device time in milliseconds (CUDA events on the handle's stream) */
double jwas_last_sweep_ms(jwas_handle* h);
/* device time (ms, CUDA events on the handle's stream) and launch count of the dominant
 * genotype-streaming kernel(s) of the last sweep; needs jwas_set_option(h,"profile",1) */
double jwas_last_stream_kernel_ms(jwas_handle* h, int64_t* launches);
/* engine 1 phase timers of the last sweep, nanoseconds: out[0..4] = CTA 0 {wait previous chain,
 * axpy+quantise+tables, stream, wait for all slices, chain}; out[8..12] = CTA 1, same phases;
 * out[16..20] = the dedicated chain CTA of the lagged schedule (option "lag" = 1);
 * out[24..28] = inside the chain {preload issue, wait, rhs + corrections, rounds, epilogue} (32 values) */
int jwas_get_phase_ns(jwas_handle* h, uint64_t* out32);
/* raw CUDA stream of the handle (cudaStream_t) so callers can time on it */
void* jwas_stream(jwas_handle* h);

#ifdef __cplusplus
}
#endif
#endif
