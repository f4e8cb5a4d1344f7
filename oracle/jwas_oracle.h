/*
 * jwas_oracle.h -- CPU oracle for the JWAS marker-effects sweep.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under jwas.jl_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Parity status: the reference is Julia and Julia is not installed in this image,
 * so the oracle cannot be pinned against outputs of JWAS itself ("same-seed chain
 * equality" is PARITY UNPINNED here).  It IS pinned against every RNG-free
 * known answer the reference's own tests hold for this path (codec bit layout,
 * decode with missing=9, xpRinvx, closed-form multi-trait state probabilities,
 * BayesR sufficient statistics, block repetition schedule, block-start validation);
 * see tests/test_oracle_pins.py and tests/golden/.
 *
 * Two arithmetic flavours:
 *   *_ref      -- faithful restatement of the reference arithmetic (Float32 data,
 *                 Float32/Float64 scalar mix exactly as Julia promotes it, libm).
 *                 Used for the reference pins and as the timed CPU baseline.
 *   *_contract -- same algorithm under the B200 arithmetic contract
 *                 (include/jwas_contract.h): exact fixed-point dots, binary64 scalar
 *                 path, deterministic log/exp, Philox draws.  The CUDA kernels must
 *                 match these bit for bit.
 */
#ifndef JWAS_ORACLE_H
#define JWAS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- codec: streaming_genotypes.jl:364-367, 622-627 (pack), 978-1002 (decode) ---- */
void jwo_pack_codes(const int8_t* codes, int64_t n, int64_t p, uint8_t* packed, int64_t stride);
void jwo_decode_marker(const uint8_t* col, int64_t n, float mean, int centered, float* dest);
/* streaming_genotypes.jl:1009-1027 */
void jwo_mul_alpha(const uint8_t* packed, int64_t n, int64_t p, int64_t stride,
                   const float* means, const float* alpha, float* out);

/* ---- per-marker statistics ---- */
/* reference arithmetic: streaming_genotypes.jl:546-585 (dense converter), Float32 sequential */
void jwo_marker_stats_ref(const uint8_t* packed, int64_t n, int64_t p, int64_t stride, int center,
                          float* means, float* xpx, float* afreq);
/* contract arithmetic: integer counts -> binary64 closed form -> float */
void jwo_marker_stats(const uint8_t* packed, int64_t n, int64_t p, int64_t stride,
                      float* means, float* xpx);
/* Gram block X_b' X_b (tools4genotypes.jl:263), contract arithmetic, row-major b*b */
void jwo_gram_block(const uint8_t* packed, int64_t n, int64_t stride, const float* means,
                    int64_t j0, int64_t b, float* G);

/* ---- reference-arithmetic samplers (dense Float32, draws replayed from u[], z[]) ---- */
/* BayesC0L.jl:25-47 BayesL! ; ngamma == 1 is BayesC0! (RR-BLUP, :19-23) */
void jwo_bayesl_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                    float* ycorr, float* alpha, const double* gamma, int64_t ngamma,
                    float vRes, float vEff, const double* z, int nthreads);

/* MTBayesC0L.jl:11-58 MTBayesL! ; ngamma == 1 is MTBayesC0! (multi-trait RR-BLUP).  alpha (t, p), ycorr (t*n) */
void jwo_mtbayesl_ref(const float* X, int64_t n, int64_t p, int t, const float* xpx,
                      float* ycorr, float* alpha, const double* gamma, int64_t ngamma,
                      const double* R, const double* G, const double* z);

/* BayesABC.jl:24-80.  nthreads parallelises the n-long dot/axpy only (BLAS threads). */
void jwo_bayesabc_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                      float* ycorr, float* alpha, float* beta, float* delta,
                      float vare, const float* varEffects, const double* pi,
                      const double* u, const double* z, int nthreads);
/* BayesABC.jl:88-108 (decode_marker! + the same update) */
void jwo_bayesabc_streaming_ref(const uint8_t* packed, int64_t n, int64_t p, int64_t stride,
                                const float* means, const float* xpx,
                                float* ycorr, float* alpha, float* beta, float* delta,
                                float vare, const float* varEffects, const double* pi,
                                const double* u, const double* z);
/* BayesABC.jl:118-188 (exact) and :190-255 (independent); nreps<=0 -> block size.
 * u,z indexed [rep*p + j]. */
void jwo_bayesabc_block_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                            const int64_t* starts, int64_t nblocks, int nreps, int independent,
                            float* ycorr, float* alpha, float* beta, float* delta,
                            float vare, const float* varEffects, const double* pi,
                            const double* u, const double* z);
/* BayesR.jl:45-97 ; pi4 has nclasses entries (or p*nclasses row-major if per_marker_pi) */
void jwo_bayesr_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                    float* ycorr, float* alpha, int32_t* delta,
                    float vare, float sigmaSq, const double* pi, int per_marker_pi,
                    const double* gamma, int nclasses,
                    const double* u, const double* z, int nthreads);
/* MTBayesABC.jl:57-127 (sampler I).  ycorr: t*n, alpha/beta/delta: t*p, R,G: t*t
 * row-major (G: per-marker p*t*t if per_marker_G), bigPi: 2^t entries indexed
 * sum(delta_k << k) (annotation_setup.jl:26-39) or p*2^t if per_marker_pi.
 * u,z indexed [k*p + j]. */
void jwo_mtbayesabc_I_ref(const float* X, int64_t n, int64_t p, int t, const float* xpx,
                          float* ycorr, float* alpha, float* beta, float* delta,
                          const double* R, const double* G, int per_marker_G,
                          const double* bigPi, int per_marker_pi,
                          const double* u, const double* z);
/* MTBayesABC.jl:129-210 (sampler II, joint 2^t states; t == 2 only here) */
void jwo_mtbayesabc_II_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                           float* ycorr, float* alpha, float* beta, float* delta,
                           const double* R, const double* G, const double* bigPi,
                           const double* u, const double* z2 /* [j*2 + k] */);

/* BayesR.jl:111-193 (exact block), :195-273 (independent); nreps <= 0 -> block size; u,z [rep*p + j] */
void jwo_bayesr_block_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                          const int64_t* starts, int64_t nblocks, int nreps, int independent,
                          float* ycorr, float* alpha, int32_t* delta,
                          float vare, float sigmaSq, const double* pi, int per_marker_pi,
                          const double* gamma, int nclasses, const double* u, const double* z);
/* MTBayesABC.jl:243-333 / :335-437 (sampler = 1) and :439-537 / :539-646 (sampler = 2, t == 2);
 * global G and Pi; u,z [(rep*t + k)*p + j] */
void jwo_mtbayesabc_block_ref(const float* X, int64_t n, int64_t p, int t, int sampler, const float* xpx,
                              const int64_t* starts, int64_t nblocks, int nreps, int independent,
                              float* ycorr, float* alpha, float* beta, float* delta,
                              const double* R, const double* G, const double* bigPi,
                              const double* u, const double* z);

/* ---- schedule helpers ---- */
int jwo_bayesr_block_nreps(int64_t iter, int64_t burnin, int64_t block_size); /* BayesR.jl:22-25 */
int jwo_validate_block_starts(const int64_t* starts, int64_t nstarts, int64_t nmarkers); /* JWAS.jl:73-79; 0 ok */
/* variance_components.jl:68-79 */
void jwo_bayesr_sigma_sufficient_statistics(const float* alpha, const int32_t* delta,
                                            const double* gamma, int64_t p, double* ssq, int64_t* nnz);

/* BayesB per-marker variances under the contract stream (variance_components.jl:169-172) */
void jwo_bayesb_variances(const float* beta, int64_t p, double df, double scale,
                          uint64_t seed, uint32_t iter, double* ve);

/* ---- contract-arithmetic sweep (what the CUDA path must reproduce bit for bit) ---- */
#define JWO_METHOD_ABC 0   /* BayesA/B/C */
#define JWO_METHOD_R   1   /* BayesR     */
#define JWO_METHOD_MT1 2   /* multi-trait BayesABC sampler I */
#define JWO_METHOD_MT2 3   /* multi-trait BayesABC sampler II (joint 2^t states), t = 2 */
#define JWO_METHOD_MEGA 4  /* megaBayesABC! (constraint=true): one single-trait BayesABC step per trait;
                              vare, varEffects, pi come per trait through Rmat / Gmat diagonals and bigPi[k] */

typedef struct {
    /* genotypes */
    int64_t n, p, stride;
    const uint8_t* packed;
    const float* means;
    const float* xpx;
    /* schedule: blocks [starts[i], starts[i+1]) ; nreps_mode 0 -> 1 rep, 1 -> block size */
    const int64_t* starts; int64_t nblocks;
    int nreps_mode; int independent;
    /* method */
    int method; int ntraits;
    /* state (in/out) */
    float* ycorr;       /* ntraits*n */
    float* alpha;       /* ntraits*p */
    float* beta;        /* ntraits*p (ABC, MT1) */
    int32_t* delta;     /* ntraits*p (0/1, BayesR 1..4) */
    /* hyper-parameters */
    double vare;                 /* ABC, R */
    const double* varEffects;    /* ABC: p entries */
    const double* pi;            /* ABC: p entries (P(effect==0)); R: nclasses or p*nclasses */
    int per_marker_pi;
    double sigmaSq; const double* gamma; int nclasses;      /* R */
    const double* Rmat; const double* Gmat; int per_marker_G; /* MT1: t*t row-major */
    const double* bigPi;                                      /* MT1: 2^t or p*2^t */
    /* draws */
    uint64_t seed; uint32_t iter;
    const double* u; const double* z;  /* replay tables [ (rep*ntraits + trait)*p + j ] or NULL */
    /* outputs */
    int overflow;                /* sticky: fixed-point clamp was hit */
    int scale_exp;               /* S used */
    /* row-sharded emulation (multi-GPU host logic on CPU): this "rank" owns rows [row_begin,row_end)
     * (row_end == 0 -> all rows); allreduce sums an int64 buffer in place over the ranks */
    int64_t row_begin, row_end;
    void (*allreduce)(void* ctx, int64_t* buf, int64_t count);
    void* ctx;
    /* lag = L in 1..3 (exact schedule only): block k's rhs is computed from ycorr carrying the updates of
     * blocks <= k-L-1 and corrected by the cross-Gram of the updates of blocks k-L .. k-1,
     *   r_j += sum_{a in blocks k-L..k-1, oldest block first, commit order} d_a * x_a'x_j,
     * so that the chains of the L previous blocks can overlap the streaming of block k on the device.  Same
     * chain in exact arithmetic; differs from lag = 0 in rounding only. */
    int lag;
} jwo_sweep_args;

int jwo_sweep_contract(jwo_sweep_args* a);

/* number of host threads the *_ref samplers will use for nthreads<=0 */
int jwo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
