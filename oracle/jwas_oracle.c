/*
 * jwas_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see jwas_oracle.h).
 *
 * Every function cites the reference lines it restates; paths are relative to
 * /root/reference/src/1.JWAS/src/.  Nothing here is copied: the reference is
 * Julia, this is a from-scratch C restatement of the arithmetic.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off; the dense dot/axpy live in
 * jwas_oracle_blas.c which uses -O3 -march=x86-64-v3 -fopenmp).
 */
#include "jwas_oracle.h"
#include "../include/jwas_contract.h"
#include <stdlib.h>
#include <stdio.h>
#include <math.h>

/* dense level-1 kernels (jwas_oracle_blas.c): stand-ins for OpenBLAS sdot/saxpy */
float jwo_sdot(const float* x, const float* y, int64_t n, int nthreads);
void  jwo_saxpy(float a, const float* x, float* y, int64_t n, int nthreads);
int   jwo_max_threads(void);

/* ======================================================================== */
/* codec                                                                    */
/* ======================================================================== */

/* markers/streaming_genotypes.jl:622-627: code = missing ? 3 : round(value);
 * byte (i-1)>>2, shift ((i-1)&3)<<1, marker-major with stride cld(nObs,4). */
void jwo_pack_codes(const int8_t* codes, int64_t n, int64_t p, uint8_t* packed, int64_t stride) {
    for (int64_t j = 0; j < p; ++j) {
        uint8_t* col = packed + j * stride;
        for (int64_t b = 0; b < stride; ++b) col[b] = 0;
        for (int64_t i = 0; i < n; ++i) {
            int8_t c = codes[j * n + i];
            uint8_t code = (c == 0 || c == 1 || c == 2) ? (uint8_t)c : (uint8_t)3;
            col[i >> 2] |= (uint8_t)(code << ((i & 3) << 1));
        }
    }
}

/* markers/streaming_genotypes.jl:978-1002 decode_marker! */
void jwo_decode_marker(const uint8_t* col, int64_t n, float mean, int centered, float* dest) {
    for (int64_t i = 0; i < n; ++i) {
        unsigned code = (col[i >> 2] >> ((i & 3) << 1)) & 3u;
        float v = (code == 3u) ? mean : (float)code;
        dest[i] = centered ? (v - mean) : v;
    }
}

/* markers/streaming_genotypes.jl:1009-1027 streaming_mul_alpha! */
void jwo_mul_alpha(const uint8_t* packed, int64_t n, int64_t p, int64_t stride,
                   const float* means, const float* alpha, float* out) {
    float* buf = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) out[i] = 0.0f;
    for (int64_t j = 0; j < p; ++j) {
        if (alpha[j] != 0.0f) {
            jwo_decode_marker(packed + j * stride, n, means[j], 1, buf);
            jwo_saxpy(alpha[j], buf, out, n, 1);
        }
    }
    free(buf);
}

/* ======================================================================== */
/* per-marker statistics                                                    */
/* ======================================================================== */

/* markers/streaming_genotypes.jl:546-585 (dense converter): Float32, in index order */
void jwo_marker_stats_ref(const uint8_t* packed, int64_t n, int64_t p, int64_t stride, int center,
                          float* means, float* xpx, float* afreq) {
    for (int64_t j = 0; j < p; ++j) {
        const uint8_t* col = packed + j * stride;
        int64_t nn = 0; float sum = 0.0f;
        for (int64_t i = 0; i < n; ++i) {
            unsigned c = jw_code(col, i);
            if (c != 3u) { nn += 1; sum += (float)c; }
        }
        float mu = sum / (float)nn;
        means[j] = mu;
        if (afreq) afreq[j] = mu / 2.0f;
        float ssc = 0.0f, ssr = 0.0f;
        for (int64_t i = 0; i < n; ++i) {
            unsigned c = jw_code(col, i);
            float v = (c == 3u) ? mu : (float)c;
            float d = v - mu;
            ssc += d * d;
            ssr += v * v;
        }
        xpx[j] = center ? ssc : ssr;
    }
}

/* integer sufficient statistics of a marker pair; everything the centred
 * cross-product needs when either marker has missing calls. */
typedef struct { int64_t Nab, Sa_vb, Sb_va, Nvv; } jwo_pair;

/* The four sums for one pair of packed bytes (four individuals each), 16 bits per sum:
 * bits 0-15 Nab, 16-31 Sa_vb, 32-47 Sb_va, 48-63 Nvv.  Pure integer bookkeeping: the counts are the ones the
 * per-individual loop gives, only four individuals at a time (this keeps the oracle side of the large parity
 * cases at seconds instead of minutes). */
static uint64_t pair_lut[256][256];
static int pair_lut_ready = 0;

static void pair_lut_init(void) {
    for (int x = 0; x < 256; ++x)
        for (int y = 0; y < 256; ++y) {
            uint64_t nab = 0, sa = 0, sb = 0, nvv = 0;
            for (int k = 0; k < 4; ++k) {
                unsigned a = (unsigned)(x >> (2 * k)) & 3u, b = (unsigned)(y >> (2 * k)) & 3u;
                if (a != 3u && b != 3u) { nab += a * b; sa += a; sb += b; nvv += 1; }
            }
            pair_lut[x][y] = nab | (sa << 16) | (sb << 32) | (nvv << 48);
        }
    pair_lut_ready = 1;
}

static jwo_pair pair_counts(const uint8_t* ca, const uint8_t* cb, int64_t n) {
    jwo_pair q = {0, 0, 0, 0};
    if (!pair_lut_ready) pair_lut_init();      /* idempotent: a second initialiser writes the same values */
    const int64_t full = n >> 2;                 /* whole bytes; the last partial byte goes individual by individual */
    int64_t bpos = 0;
    while (bpos < full) {
        int64_t end = bpos + 4000 < full ? bpos + 4000 : full;      /* 4000 bytes * 16 < 65536: no field overflows */
        uint64_t acc = 0;
        for (; bpos < end; ++bpos) acc += pair_lut[ca[bpos]][cb[bpos]];
        q.Nab += (int64_t)(acc & 0xffffu); q.Sa_vb += (int64_t)((acc >> 16) & 0xffffu);
        q.Sb_va += (int64_t)((acc >> 32) & 0xffffu); q.Nvv += (int64_t)(acc >> 48);
    }
    for (int64_t i = full << 2; i < n; ++i) {
        unsigned a = jw_code(ca, i), b = jw_code(cb, i);
        if (a != 3u && b != 3u) { q.Nab += (int64_t)(a * b); q.Sa_vb += a; q.Sb_va += b; q.Nvv += 1; }
    }
    return q;
}

/* Contract definition of x_a' x_b for centred, mean-imputed columns
 * (x = code - mu where observed, 0 where missing; readgenotypes.jl:372-385):
 *   sum_{valid both} (a-mu_a)(b-mu_b) = Nab - mu_a*Sb - mu_b*Sa + mu_a*mu_b*Nvv
 * evaluated in binary64 in exactly this order, then rounded to Float32. */
static float gram_value(jwo_pair q, float mua, float mub) {
    double ma = (double)mua, mb = (double)mub;
    double g = (double)q.Nab - ma * (double)q.Sb_va;
    g = g - mb * (double)q.Sa_vb;
    g = g + (ma * mb) * (double)q.Nvv;
    return (float)g;
}

/* contract: mean = Float32(sum)/Float32(nn) (streaming_genotypes.jl:566), xpx = Gram diagonal */
void jwo_marker_stats(const uint8_t* packed, int64_t n, int64_t p, int64_t stride,
                      float* means, float* xpx) {
    for (int64_t j = 0; j < p; ++j) {
        const uint8_t* col = packed + j * stride;
        int64_t nn = 0, sum = 0;
        for (int64_t i = 0; i < n; ++i) {
            unsigned c = jw_code(col, i);
            if (c != 3u) { nn += 1; sum += c; }
        }
        float mu = (nn > 0) ? (float)sum / (float)nn : 0.0f;
        means[j] = mu;
        jwo_pair q = pair_counts(col, col, n);
        xpx[j] = gram_value(q, mu, mu);
    }
}

/* tools4genotypes.jl:263 XpRinvX = Xblock' * Xblock (unit weights) */
void jwo_gram_block(const uint8_t* packed, int64_t n, int64_t stride, const float* means,
                    int64_t j0, int64_t b, float* G) {
    for (int64_t a = 0; a < b; ++a)
        for (int64_t c = a; c < b; ++c) {
            jwo_pair q = pair_counts(packed + (j0 + a) * stride, packed + (j0 + c) * stride, n);
            float g = gram_value(q, means[j0 + a], means[j0 + c]);
            G[a * b + c] = g;
            if (c != a) {
                /* the (c,a) entry is evaluated with the roles swapped so that the
                 * stored matrix is exactly what a kernel computing row c would get */
                jwo_pair r = { q.Nab, q.Sb_va, q.Sa_vb, q.Nvv };
                G[c * b + a] = gram_value(r, means[j0 + c], means[j0 + a]);
            }
        }
}

/* ======================================================================== */
/* reference-arithmetic samplers                                            */
/* ======================================================================== */

/* BayesABC.jl:24-58 bayesabc_update_marker!, with Julia's promotion rules:
 * Float32 data and residual/effect variances, Float64 pi and random draws. */
static void abc_update_marker_ref(const float* x, int64_t n, float* yCorr,
                                  float* alpha, float* beta, float* delta, int64_t j,
                                  float xRinvy, float xpx_j, float invVarRes,
                                  float invVarEffect_j, float logVarEffect_j, float varEffect_j,
                                  double logDelta0, double logPiComp,
                                  double u, double z, int nthreads) {
    float rhs = (xRinvy + xpx_j * alpha[j]) * invVarRes;             /* :36 */
    float lhs = xpx_j * invVarRes + invVarEffect_j;                  /* :37 */
    float invLhs = 1.0f / lhs;                                       /* :38 */
    float gHat = rhs * invLhs;                                       /* :39 */
    float inner = logf(lhs) + logVarEffect_j - gHat * rhs;
    double logDelta1 = -0.5 * (double)inner + logPiComp;             /* :40 */
    double probDelta1 = 1.0 / (1.0 + exp(logDelta0 - logDelta1));    /* :41 */
    float oldAlpha = alpha[j];
    if (u < probDelta1) {                                            /* :44 */
        delta[j] = 1.0f;
        beta[j] = (float)((double)gHat + z * (double)sqrtf(invLhs)); /* :46 */
        alpha[j] = beta[j];
        jwo_saxpy(oldAlpha - alpha[j], x, yCorr, n, nthreads);       /* :48 */
    } else {
        if (oldAlpha != 0.0f) jwo_saxpy(oldAlpha, x, yCorr, n, nthreads); /* :50-52 */
        delta[j] = 0.0f;
        beta[j] = (float)(z * (double)sqrtf(varEffect_j));           /* :54 */
        alpha[j] = 0.0f;
    }
}

/* BayesABC.jl:60-80 BayesABC! */
void jwo_bayesabc_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                      float* ycorr, float* alpha, float* beta, float* delta,
                      float vare, const float* varEffects, const double* pi,
                      const double* u, const double* z, int nthreads) {
    float invVarRes = 1.0f / vare;
    for (int64_t j = 0; j < p; ++j) {
        const float* x = X + j * n;
        float dot = jwo_sdot(x, ycorr, n, nthreads);                 /* :76 */
        abc_update_marker_ref(x, n, ycorr, alpha, beta, delta, j, dot, xpx[j], invVarRes,
                              1.0f / varEffects[j], logf(varEffects[j]), varEffects[j],
                              log(pi[j]), log(1.0 - pi[j]), u[j], z[j], nthreads);
    }
}

/* BayesC0L.jl:25-47 BayesL! (and BayesC0! = the same routine with gammaArray = [1.0], :19-23).
 * Julia's promotion: xpRinvx, alpha, yCorr, vRes, vEff Float32; gammaArray Float64 (rand(Gamma(1,8)),
 * MCMC_BayesianAlphabet.jl:72-77); the literal 1.0 makes invLhs Float64 either way. */
void jwo_bayesl_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                    float* ycorr, float* alpha, const double* gamma, int64_t ngamma,
                    float vRes, float vEff, const double* z, int nthreads) {
    float lambda = vRes / vEff;                                      /* :30 */
    for (int64_t j = 0; j < p; ++j) {
        const float* x = X + j * n;
        float rhs = jwo_sdot(x, ycorr, n, nthreads) + xpx[j] * alpha[j];   /* :39 */
        double invLhs;
        if (ngamma > 1) invLhs = 1.0 / ((double)xpx[j] + (double)lambda / gamma[j]);   /* :32, :40-41 */
        else invLhs = 1.0 / (double)(xpx[j] + lambda);               /* :33: Float32 sum, then 1.0/ */
        double mean = invLhs * (double)rhs;                          /* :42 */
        float oldAlpha = alpha[j];
        alpha[j] = (float)(mean + z[j] * sqrt(invLhs * (double)vRes));   /* :44 */
        jwo_saxpy(oldAlpha - alpha[j], x, ycorr, n, nthreads);       /* :45 */
    }
}

/* BayesABC.jl:88-108 BayesABC_streaming! */
void jwo_bayesabc_streaming_ref(const uint8_t* packed, int64_t n, int64_t p, int64_t stride,
                                const float* means, const float* xpx,
                                float* ycorr, float* alpha, float* beta, float* delta,
                                float vare, const float* varEffects, const double* pi,
                                const double* u, const double* z) {
    float invVarRes = 1.0f / vare;
    float* buf = (float*)malloc(sizeof(float) * (size_t)n);
    for (int64_t j = 0; j < p; ++j) {
        jwo_decode_marker(packed + j * stride, n, means[j], 1, buf);  /* :101 */
        float dot = jwo_sdot(buf, ycorr, n, 1);
        abc_update_marker_ref(buf, n, ycorr, alpha, beta, delta, j, dot, xpx[j], invVarRes,
                              1.0f / varEffects[j], logf(varEffects[j]), varEffects[j],
                              log(pi[j]), log(1.0 - pi[j]), u[j], z[j], 1);
    }
    free(buf);
}

/* BayesABC.jl:118-188 (exact) / :190-255 (independent).  Gram in Float32
 * (tools4genotypes.jl:263 sgemm), block rhs by sgemv (tools4genotypes.jl:64-66). */
void jwo_bayesabc_block_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                            const int64_t* starts, int64_t nblocks, int nreps_in, int independent,
                            float* ycorr, float* alpha, float* beta, float* delta,
                            float vare, const float* varEffects, const double* pi,
                            const double* u, const double* z) {
    float invVarRes = 1.0f / vare;
    float* snap = NULL;
    float* dsave = NULL;
    if (independent) {
        snap = (float*)malloc(sizeof(float) * (size_t)n);            /* :205 */
        for (int64_t i = 0; i < n; ++i) snap[i] = ycorr[i];
        dsave = (float*)malloc(sizeof(float) * (size_t)p);
    }
    for (int64_t ib = 0; ib < nblocks; ++ib) {
        int64_t s = starts[ib], e = starts[ib + 1], b = e - s;
        float* G = (float*)malloc(sizeof(float) * (size_t)(b * b));
        float* r = (float*)malloc(sizeof(float) * (size_t)b);
        float* aold = (float*)malloc(sizeof(float) * (size_t)b);
        for (int64_t a = 0; a < b; ++a)
            for (int64_t c = 0; c < b; ++c)
                G[a * b + c] = jwo_sdot(X + (s + a) * n, X + (s + c) * n, n, 1);
        for (int64_t a = 0; a < b; ++a) {
            aold[a] = alpha[s + a];                                  /* :150 */
            r[a] = jwo_sdot(X + (s + a) * n, independent ? snap : ycorr, n, 1); /* :152 / :217 */
        }
        int nreps = nreps_in > 0 ? nreps_in : (int)b;                /* :153 */
        for (int rep = 0; rep < nreps; ++rep) {
            for (int64_t jj = 0; jj < b; ++jj) {
                int64_t j = s + jj;
                float ve = varEffects[j];
                float rhs = (r[jj] + xpx[j] * alpha[j]) * invVarRes; /* :157 */
                float lhs = xpx[j] * invVarRes + 1.0f / ve;
                float invLhs = 1.0f / lhs;
                float gHat = rhs * invLhs;
                float inner = logf(lhs) + logf(ve) - gHat * rhs;
                double logDelta1 = -0.5 * (double)inner + log(1.0 - pi[j]);
                double prob = 1.0 / (1.0 + exp(log(pi[j]) - logDelta1));
                float oldAlpha = alpha[j];
                double uu = u[(int64_t)rep * p + j], zz = z[(int64_t)rep * p + j];
                float a;
                if (uu < prob) {
                    delta[j] = 1.0f;
                    beta[j] = (float)((double)gHat + zz * (double)sqrtf(invLhs));
                    alpha[j] = beta[j];
                    a = oldAlpha - alpha[j];
                    for (int64_t m = 0; m < b; ++m) r[m] += a * G[m * b + jj]; /* :169 */
                } else {
                    if (oldAlpha != 0.0f) {
                        a = oldAlpha;
                        for (int64_t m = 0; m < b; ++m) r[m] += a * G[m * b + jj]; /* :172 */
                    }
                    delta[j] = 0.0f;
                    beta[j] = (float)(zz * (double)sqrtf(ve));
                    alpha[j] = 0.0f;
                }
            }
        }
        for (int64_t a = 0; a < b; ++a) aold[a] -= alpha[s + a];     /* :183 */
        if (independent) {
            for (int64_t a = 0; a < b; ++a) dsave[s + a] = aold[a];  /* :247-248 */
        } else {
            for (int64_t a = 0; a < b; ++a)                          /* :184 mul!(yCorr,X_b,d,1,1) */
                if (aold[a] != 0.0f) jwo_saxpy(aold[a], X + (s + a) * n, ycorr, n, 1);
        }
        free(G); free(r); free(aold);
    }
    if (independent) {                                               /* :251-253 */
        for (int64_t j = 0; j < p; ++j)
            if (dsave[j] != 0.0f) jwo_saxpy(dsave[j], X + j * n, ycorr, n, 1);
        free(snap); free(dsave);
    }
}

/* Distributions.jl 0.25 rand(Categorical(p)): first i with cumsum > u, last class as fallback */
static int categorical_from_uniform(const double* probs, int k, double u) {
    double cp = probs[0]; int i = 0;
    while (cp <= u && i < k - 1) { i += 1; cp += probs[i]; }
    return i;
}

/* BayesR.jl:45-97 BayesR! */
void jwo_bayesr_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                    float* ycorr, float* alpha, int32_t* delta,
                    float vare, float sigmaSq, const double* pi, int per_marker_pi,
                    const double* gamma, int nclasses,
                    const double* u, const double* z, int nthreads) {
    double log_probs[16], probs[16];
    float invVarRes = 1.0f / vare;                                   /* :55 */
    for (int64_t j = 0; j < p; ++j) {
        const float* x = X + j * n;
        float rhs = (jwo_sdot(x, ycorr, n, nthreads) + xpx[j] * alpha[j]) * invVarRes; /* :60 */
        float oldAlpha = alpha[j];
        const double* pij = per_marker_pi ? pi + j * nclasses : pi;
        log_probs[0] = log(pij[0]);                                  /* :64 */
        for (int k = 1; k < nclasses; ++k) {
            double varEffect = gamma[k] * (double)sigmaSq;           /* :66 */
            double invVarEffect = 1.0 / varEffect;
            double lhs = (double)(xpx[j] * invVarRes) + invVarEffect;
            double invLhs = 1.0 / lhs;
            double betaHat = invLhs * (double)rhs;
            log_probs[k] = 0.5 * (log(invLhs) - log(varEffect) + betaHat * (double)rhs) + log(pij[k]);
        }
        double mx = log_probs[0];
        for (int k = 1; k < nclasses; ++k) if (log_probs[k] > mx) mx = log_probs[k];
        double se = 0.0;
        for (int k = 0; k < nclasses; ++k) se += exp(log_probs[k] - mx);
        double log_norm = mx + log(se);                              /* BayesR.jl:1-4 */
        for (int k = 0; k < nclasses; ++k) probs[k] = exp(log_probs[k] - log_norm);
        int cls = categorical_from_uniform(probs, nclasses, u[j]);   /* :79 */
        delta[j] = cls + 1;
        if (cls == 0) {
            if (oldAlpha != 0.0f) jwo_saxpy(oldAlpha, x, ycorr, n, nthreads); /* :83-85 */
            alpha[j] = 0.0f;
        } else {
            double varEffect = gamma[cls] * (double)sigmaSq;
            double lhs = (double)(xpx[j] * invVarRes) + 1.0 / varEffect;
            double invLhs = 1.0 / lhs;
            double betaHat = invLhs * (double)rhs;
            alpha[j] = (float)(betaHat + z[j] * sqrt(invLhs));       /* :93 */
            jwo_saxpy(oldAlpha - alpha[j], x, ycorr, n, nthreads);   /* :94 */
        }
    }
}

static void inv_small(const double* A, int t, double* Ai) {
    /* Gauss-Jordan with partial pivoting, t <= 8 */
    double M[8][16];
    for (int i = 0; i < t; ++i) {
        for (int j = 0; j < t; ++j) { M[i][j] = A[i * t + j]; M[i][t + j] = (i == j) ? 1.0 : 0.0; }
    }
    for (int c = 0; c < t; ++c) {
        int piv = c;
        for (int r = c + 1; r < t; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 2 * t; ++j) { double tmp = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = tmp; }
        double d = M[c][c];
        for (int j = 0; j < 2 * t; ++j) M[c][j] /= d;
        for (int r = 0; r < t; ++r) if (r != c) {
            double f = M[r][c];
            if (f != 0.0) for (int j = 0; j < 2 * t; ++j) M[r][j] -= f * M[c][j];
        }
    }
    for (int i = 0; i < t; ++i) for (int j = 0; j < t; ++j) Ai[i * t + j] = M[i][t + j];
}

/* MTBayesC0L.jl:11-58 MTBayesL! ; ngamma == 1 is MTBayesC0! (multi-trait RR-BLUP, :6-9).
 * Every marker stays in the model: per marker the traits are drawn one after the other from their full conditionals
 * given the other traits' current effects (:44-51), all axpys after the trait loop (:52-54).  R, G arrive as the
 * Float32-valued matrices the reference holds; their inverses are taken in binary64 here (Julia: Float32 LU). */
void jwo_mtbayesl_ref(const float* X, int64_t n, int64_t p, int t, const float* xpx,
                      float* ycorr, float* alpha, const double* gamma, int64_t ngamma,
                      const double* R, const double* G, const double* z) {
    double Rinv[64], Ginv[64];
    inv_small(R, t, Rinv);                                           /* :17 */
    inv_small(G, t, Ginv);                                           /* :18 */
    double Rhs[8], rr[8], Lhs[64];
    float newa[8], olda[8];
    for (int64_t m = 0; m < p; ++m) {
        const float* x = X + m * n;
        for (int k = 0; k < t; ++k) {                                /* :36-39 */
            olda[k] = newa[k] = alpha[k * p + m];
            Rhs[k] = (double)(jwo_sdot(x, ycorr + k * n, n, 1) + xpx[m] * olda[k]);
        }
        for (int k = 0; k < t; ++k) {                                /* :40 Rhs = invR0*Rhs */
            rr[k] = 0.0;
            for (int q = 0; q < t; ++q) rr[k] += Rinv[k * t + q] * Rhs[q];
        }
        const double g = ngamma > 1 ? gamma[m] : 1.0;
        for (int q = 0; q < t * t; ++q) Lhs[q] = (double)xpx[m] * Rinv[q] + Ginv[q] / g;   /* :41, :27-30 */
        for (int k = 0; k < t; ++k) {                                /* :42-51 */
            double lhs = Lhs[k * t + k], ilhs = 1.0 / lhs, dot = 0.0;
            for (int q = 0; q < t; ++q) dot += Lhs[k * t + q] * (double)newa[q];
            double mu = ilhs * (rr[k] - dot) + (double)newa[k];
            newa[k] = (float)(mu + z[k * p + m] * sqrt(ilhs));
            alpha[k * p + m] = newa[k];
        }
        for (int k = 0; k < t; ++k) jwo_saxpy(olda[k] - newa[k], x, ycorr + k * n, n, 1);   /* :52-54 */
    }
}

/* MTBayesABC.jl:57-127 _MTBayesABC_samplerI! */
void jwo_mtbayesabc_I_ref(const float* X, int64_t n, int64_t p, int t, const float* xpx,
                          float* ycorr, float* alpha, float* beta, float* delta,
                          const double* R, const double* G, int per_marker_G,
                          const double* bigPi, int per_marker_pi,
                          const double* u, const double* z) {
    double Rinv[64], Ginv[64];
    inv_small(R, t, Rinv);                                           /* :66 */
    if (!per_marker_G) inv_small(G, t, Ginv);                        /* :67 */
    int nstates = 1 << t;
    double b[8], newa[8], olda[8], d[8], w[8];
    for (int64_t m = 0; m < p; ++m) {
        const float* x = X + m * n;
        if (per_marker_G) inv_small(G + m * t * t, t, Ginv);
        const double* Pi = per_marker_pi ? bigPi + m * nstates : bigPi;
        for (int k = 0; k < t; ++k) {                                /* :78-83 */
            b[k] = beta[k * p + m];
            olda[k] = newa[k] = alpha[k * p + m];
            d[k] = delta[k * p + m];
            w[k] = (double)(jwo_sdot(x, ycorr + k * n, n, 1) + xpx[m] * alpha[k * p + m]);
        }
        for (int k = 0; k < t; ++k) {                                /* :85-119 */
            double Ginv11 = Ginv[k * t + k];
            double C11 = Ginv11 + Rinv[k * t + k] * (double)xpx[m];
            double rhs0 = 0.0, c12b = 0.0;
            for (int q = 0; q < t; ++q) if (q != k) {
                double Ginv12 = Ginv[k * t + q];
                double C12 = Ginv12 + (double)xpx[m] * d[q] * Rinv[k * t + q]; /* :90 */
                rhs0 -= Ginv12 * b[q];                               /* :93 */
                c12b += C12 * b[q];
            }
            double invLhs0 = 1.0 / Ginv11;
            double gHat0 = rhs0 * invLhs0;
            double invLhs1 = 1.0 / C11;
            double wr = 0.0;
            for (int q = 0; q < t; ++q) wr += w[q] * Rinv[q * t + k];
            double rhs1 = wr - c12b;                                 /* :96 */
            double gHat1 = rhs1 * invLhs1;
            int s0 = 0, s1 = 0;
            for (int q = 0; q < t; ++q) {
                int dq = (q == k) ? 0 : (d[q] != 0.0);
                s0 |= dq << q; s1 |= ((q == k) ? 1 : dq) << q;
            }
            double logDelta0 = -0.5 * (log(Ginv11) - gHat0 * gHat0 * Ginv11) + log(Pi[s0]); /* :104 */
            double logDelta1 = -0.5 * (log(C11) - gHat1 * gHat1 * C11) + log(Pi[s1]);       /* :105 */
            double prob1 = 1.0 / (1.0 + exp(logDelta0 - logDelta1));
            double uu = u[k * p + m], zz = z[k * p + m];
            float* yk = ycorr + k * n;
            if (uu < prob1) {                                        /* :108 */
                d[k] = 1.0;
                float nb = (float)(gHat1 + zz * sqrt(invLhs1));
                b[k] = newa[k] = nb;
                jwo_saxpy((float)olda[k] - nb, x, yk, n, 1);         /* :111 */
            } else {
                b[k] = (double)(float)(gHat0 + zz * sqrt(invLhs0));  /* :113 */
                d[k] = 0.0;
                newa[k] = 0.0;
                if (olda[k] != 0.0) jwo_saxpy((float)olda[k], x, yk, n, 1);
            }
        }
        for (int k = 0; k < t; ++k) {                                /* :121-125 */
            beta[k * p + m] = (float)b[k];
            delta[k * p + m] = (float)d[k];
            alpha[k * p + m] = (float)newa[k];
        }
    }
}

/* MTBayesABC.jl:129-210 _MTBayesABC_samplerII!, t = 2.
 * State order (0,0),(1,0),(0,1),(1,1) (annotation_setup.jl:18). */
void jwo_mtbayesabc_II_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                           float* ycorr, float* alpha, float* beta, float* delta,
                           const double* R, const double* G, const double* bigPi,
                           const double* u, const double* z2) {
    const int t = 2;
    double Rinv[4], Ginv[4];
    inv_small(R, t, Rinv); inv_small(G, t, Ginv);
    for (int64_t m = 0; m < p; ++m) {
        const float* x = X + m * n;
        double w[2], olda[2];
        for (int k = 0; k < t; ++k) {
            olda[k] = alpha[k * p + m];
            w[k] = (double)(jwo_sdot(x, ycorr + k * n, n, 1) + xpx[m] * alpha[k * p + m]); /* :168 */
        }
        double logDelta[4], bcand[4][2];
        for (int s = 0; s < 4; ++s) {
            double D[2] = { (double)(s & 1), (double)((s >> 1) & 1) };
            double lhs[4], rhs[2], ilhs[4];
            for (int i = 0; i < 2; ++i) {
                for (int j = 0; j < 2; ++j)
                    lhs[i * 2 + j] = D[i] * Rinv[i * 2 + j] * D[j] * (double)xpx[m] + Ginv[i * 2 + j]; /* :180 */
                rhs[i] = D[i] * (Rinv[0 * 2 + i] * w[0] + Rinv[1 * 2 + i] * w[1]);                    /* :181 */
            }
            double det = lhs[0] * lhs[3] - lhs[1] * lhs[2];
            ilhs[0] = lhs[3] / det; ilhs[3] = lhs[0] / det; ilhs[1] = -lhs[1] / det; ilhs[2] = -lhs[2] / det;
            double g0 = ilhs[0] * rhs[0] + ilhs[1] * rhs[1];
            double g1 = ilhs[2] * rhs[0] + ilhs[3] * rhs[1];
            logDelta[s] = -0.5 * (log(det) - (rhs[0] * g0 + rhs[1] * g1)) + log(bigPi[s]);            /* :185 */
            double L00 = sqrt(ilhs[0]), L10 = ilhs[2] / L00;
            double L11 = sqrt(ilhs[3] - L10 * L10);
            bcand[s][0] = g0 + L00 * z2[m * 2 + 0];                                                   /* :186 */
            bcand[s][1] = g1 + L10 * z2[m * 2 + 0] + L11 * z2[m * 2 + 1];
        }
        double mx = logDelta[0];
        for (int s = 1; s < 4; ++s) if (logDelta[s] > mx) mx = logDelta[s];
        double probs[4], den = 0.0;
        for (int s = 0; s < 4; ++s) { probs[s] = exp(logDelta[s] - mx); den += probs[s]; }
        for (int s = 0; s < 4; ++s) probs[s] /= den;
        int s = categorical_from_uniform(probs, 4, u[m]);                                             /* :198 */
        for (int k = 0; k < t; ++k) {
            double dk = (double)((s >> k) & 1);
            float bk = (float)bcand[s][k];
            float na = (float)(dk * (double)bk);
            jwo_saxpy((float)olda[k] - na, x, ycorr + k * n, n, 1);                                   /* :204 */
            beta[k * p + m] = bk; delta[k * p + m] = (float)dk; alpha[k * p + m] = na;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* block restatements for BayesR and the multi-trait samplers               */
/* ------------------------------------------------------------------------ */

/* Float32 Gram block Xb'Xb (tools4genotypes.jl:263, sgemm) and block rhs Xb'y (tools4genotypes.jl:64-66, sgemv) */
static void ref_gram_f32(const float* X, int64_t n, int64_t s, int64_t b, float* G) {
    for (int64_t a = 0; a < b; ++a)
        for (int64_t c = 0; c < b; ++c)
            G[a * b + c] = jwo_sdot(X + (s + a) * n, X + (s + c) * n, n, 1);
}

/* one BayesR marker step on a block rhs entry (BayesR.jl:149-183; same arithmetic as :58-95 of the dense
 * sampler).  Returns oldAlpha - newAlpha (Float32), the coefficient of the Gram-column axpy (:182). */
static float bayesr_step_ref(float r_j, float xpx_j, float* alpha_j, int32_t* delta_j, float invVarRes,
                             float sigmaSq, const double* pij, const double* gamma, int nclasses,
                             double u, double z) {
    double log_probs[16], probs[16];
    float rhs = (r_j + xpx_j * (*alpha_j)) * invVarRes;                 /* :150 */
    float oldAlpha = *alpha_j;
    log_probs[0] = log(pij[0]);                                          /* :154 */
    for (int k = 1; k < nclasses; ++k) {
        double varEffect = gamma[k] * (double)sigmaSq;
        double lhs = (double)(xpx_j * invVarRes) + 1.0 / varEffect;
        double invLhs = 1.0 / lhs;
        double betaHat = invLhs * (double)rhs;
        log_probs[k] = 0.5 * (log(invLhs) - log(varEffect) + betaHat * (double)rhs) + log(pij[k]);
    }
    double mx = log_probs[0];
    for (int k = 1; k < nclasses; ++k) if (log_probs[k] > mx) mx = log_probs[k];
    double se = 0.0;
    for (int k = 0; k < nclasses; ++k) se += exp(log_probs[k] - mx);
    double log_norm = mx + log(se);                                      /* BayesR.jl:1-4 */
    for (int k = 0; k < nclasses; ++k) probs[k] = exp(log_probs[k] - log_norm);
    int cls = categorical_from_uniform(probs, nclasses, u);              /* :168 */
    *delta_j = cls + 1;
    if (cls == 0) *alpha_j = 0.0f;                                       /* :171-172 */
    else {
        double varEffect = gamma[cls] * (double)sigmaSq;
        double lhs = (double)(xpx_j * invVarRes) + 1.0 / varEffect;
        double invLhs = 1.0 / lhs;
        double betaHat = invLhs * (double)rhs;
        *alpha_j = (float)(betaHat + z * sqrt(invLhs));                  /* :179 */
    }
    return oldAlpha - *alpha_j;
}

/* BayesR.jl:111-193 BayesR_block! (exact) and :195-273 BayesR_block_independent!.
 * nreps_in = bayesr_block_nreps(iter, burnin, block_size) evaluated by the caller per sweep: 1 during
 * burn-in, <= 0 -> the block size afterwards (:22-25, :144).  u,z indexed [rep*p + j]. */
void jwo_bayesr_block_ref(const float* X, int64_t n, int64_t p, const float* xpx,
                          const int64_t* starts, int64_t nblocks, int nreps_in, int independent,
                          float* ycorr, float* alpha, int32_t* delta,
                          float vare, float sigmaSq, const double* pi, int per_marker_pi,
                          const double* gamma, int nclasses, const double* u, const double* z) {
    float invVarRes = 1.0f / vare;                                       /* :130 */
    float* snap = NULL; float* dsave = NULL;
    if (independent) {
        snap = (float*)malloc(sizeof(float) * (size_t)n);                /* :207 */
        for (int64_t i = 0; i < n; ++i) snap[i] = ycorr[i];
        dsave = (float*)malloc(sizeof(float) * (size_t)p);
    }
    for (int64_t ib = 0; ib < nblocks; ++ib) {
        int64_t s = starts[ib], b = starts[ib + 1] - s;
        float* G = (float*)malloc(sizeof(float) * (size_t)(b * b));
        float* r = (float*)malloc(sizeof(float) * (size_t)b);
        float* aold = (float*)malloc(sizeof(float) * (size_t)b);
        ref_gram_f32(X, n, s, b, G);
        for (int64_t a = 0; a < b; ++a) {
            aold[a] = alpha[s + a];                                      /* :141 */
            r[a] = jwo_sdot(X + (s + a) * n, independent ? snap : ycorr, n, 1);   /* :143 / :222 */
        }
        int nreps = nreps_in > 0 ? nreps_in : (int)b;                    /* :144 */
        for (int rep = 0; rep < nreps; ++rep)
            for (int64_t jj = 0; jj < b; ++jj) {
                int64_t j = s + jj;
                const double* pij = per_marker_pi ? pi + j * nclasses : pi;
                float d = bayesr_step_ref(r[jj], xpx[j], &alpha[j], &delta[j], invVarRes, sigmaSq, pij, gamma,
                                          nclasses, u[(int64_t)rep * p + j], z[(int64_t)rep * p + j]);
                for (int64_t m = 0; m < b; ++m) r[m] += d * G[m * b + jj];        /* :182 (always) */
            }
        for (int64_t a = 0; a < b; ++a) aold[a] -= alpha[s + a];         /* :188 */
        if (independent) for (int64_t a = 0; a < b; ++a) dsave[s + a] = aold[a];   /* :263-264 */
        else for (int64_t a = 0; a < b; ++a)                             /* :189 mul!(yCorr, X_b, d, 1, 1) */
            if (aold[a] != 0.0f) jwo_saxpy(aold[a], X + (s + a) * n, ycorr, n, 1);
        free(G); free(r); free(aold);
    }
    if (independent) {                                                   /* :267-270 */
        for (int64_t j = 0; j < p; ++j)
            if (dsave[j] != 0.0f) jwo_saxpy(dsave[j], X + j * n, ycorr, n, 1);
        free(snap); free(dsave);
    }
}

/* MTBayesABC.jl:243-333 (block sampler I), :335-437 (independent I), :439-537 (block sampler II, t = 2),
 * :539-646 (independent II).  Same per-marker arithmetic as jwo_mtbayesabc_I_ref / _II_ref with the
 * dots replaced by the block rhs and the axpys by Gram-column updates of it (:309, :316, :524).
 * nreps_in <= 0 -> block size (:272, :486).  u,z indexed [(rep*t + k)*p + j]; sampler II uses u of
 * trait 0 for the state label and z of both traits as the shared normals (:499, :518). */
void jwo_mtbayesabc_block_ref(const float* X, int64_t n, int64_t p, int t, int sampler, const float* xpx,
                              const int64_t* starts, int64_t nblocks, int nreps_in, int independent,
                              float* ycorr, float* alpha, float* beta, float* delta,
                              const double* R, const double* G, const double* bigPi,
                              const double* u, const double* z) {
    double Rinv[64], Ginv[64];
    inv_small(R, t, Rinv); inv_small(G, t, Ginv);
    float* snap = NULL; float* dsave = NULL;
    if (independent) {
        snap = (float*)malloc(sizeof(float) * (size_t)(n * t));          /* :350 / :563 */
        for (int64_t i = 0; i < n * t; ++i) snap[i] = ycorr[i];
        dsave = (float*)calloc((size_t)(p * t), sizeof(float));
    }
    for (int64_t ib = 0; ib < nblocks; ++ib) {
        int64_t s = starts[ib], b = starts[ib + 1] - s;
        float* Gm = (float*)malloc(sizeof(float) * (size_t)(b * b));
        float* r = (float*)malloc(sizeof(float) * (size_t)(b * t));
        float* aold = (float*)malloc(sizeof(float) * (size_t)(b * t));
        ref_gram_f32(X, n, s, b, Gm);
        for (int k = 0; k < t; ++k)
            for (int64_t a = 0; a < b; ++a) {
                aold[k * b + a] = alpha[k * p + s + a];                  /* :270 */
                r[k * b + a] = jwo_sdot(X + (s + a) * n, (independent ? snap : ycorr) + k * n, n, 1);  /* :268 */
            }
        int nreps = nreps_in > 0 ? nreps_in : (int)b;
        for (int rep = 0; rep < nreps; ++rep)
            for (int64_t jj = 0; jj < b; ++jj) {
                int64_t m = s + jj;
                const double* uu = u + (int64_t)rep * t * p;
                const double* zz = z + (int64_t)rep * t * p;
                if (sampler == 2) {
                    double w[2], olda[2];
                    for (int k = 0; k < 2; ++k) {
                        olda[k] = alpha[k * p + m];
                        w[k] = (double)(r[k * b + jj] + xpx[m] * alpha[k * p + m]);              /* :496 */
                    }
                    double logDelta[4], bcand[4][2];
                    for (int st = 0; st < 4; ++st) {
                        double D[2] = { (double)(st & 1), (double)((st >> 1) & 1) };
                        double lhs[4], rhs[2], ilhs[4];
                        for (int i = 0; i < 2; ++i) {
                            for (int j = 0; j < 2; ++j)
                                lhs[i * 2 + j] = D[i] * Rinv[i * 2 + j] * D[j] * (double)xpx[m] + Ginv[i * 2 + j]; /* :501 */
                            rhs[i] = D[i] * (Rinv[0 * 2 + i] * w[0] + Rinv[1 * 2 + i] * w[1]);                    /* :502 */
                        }
                        double det = lhs[0] * lhs[3] - lhs[1] * lhs[2];
                        ilhs[0] = lhs[3] / det; ilhs[3] = lhs[0] / det; ilhs[1] = -lhs[1] / det; ilhs[2] = -lhs[2] / det;
                        double g0 = ilhs[0] * rhs[0] + ilhs[1] * rhs[1];
                        double g1 = ilhs[2] * rhs[0] + ilhs[3] * rhs[1];
                        logDelta[st] = -0.5 * (log(det) - (rhs[0] * g0 + rhs[1] * g1)) + log(bigPi[st]);          /* :506 */
                        double L00 = sqrt(ilhs[0]), L10 = ilhs[2] / L00;
                        double L11 = sqrt(ilhs[3] - L10 * L10);
                        bcand[st][0] = g0 + L00 * zz[m];                                                           /* :507 */
                        bcand[st][1] = g1 + L10 * zz[m] + L11 * zz[p + m];
                    }
                    double mx = logDelta[0];
                    for (int st = 1; st < 4; ++st) if (logDelta[st] > mx) mx = logDelta[st];
                    double probs[4], den = 0.0;
                    for (int st = 0; st < 4; ++st) { probs[st] = exp(logDelta[st] - mx); den += probs[st]; }
                    for (int st = 0; st < 4; ++st) probs[st] /= den;
                    int st = categorical_from_uniform(probs, 4, uu[m]);                                            /* :518 */
                    for (int k = 0; k < 2; ++k) {
                        double dk = (double)((st >> k) & 1);
                        float bk = (float)bcand[st][k];
                        float na = (float)(dk * (double)bk);
                        float d = (float)olda[k] - na;
                        for (int64_t q = 0; q < b; ++q) r[k * b + q] += d * Gm[q * b + jj];                        /* :524 */
                        beta[k * p + m] = bk; delta[k * p + m] = (float)dk; alpha[k * p + m] = na;
                    }
                } else {
                    double bb[8], newa[8], olda[8], d[8], w[8];
                    for (int k = 0; k < t; ++k) {                        /* :276-281 */
                        bb[k] = beta[k * p + m];
                        olda[k] = newa[k] = alpha[k * p + m];
                        d[k] = delta[k * p + m];
                        w[k] = (double)(r[k * b + jj] + xpx[m] * alpha[k * p + m]);
                    }
                    for (int k = 0; k < t; ++k) {                        /* :282-320 */
                        double Ginv11 = Ginv[k * t + k];
                        double C11 = Ginv11 + Rinv[k * t + k] * (double)xpx[m];
                        double rhs0 = 0.0, c12b = 0.0;
                        for (int q = 0; q < t; ++q) if (q != k) {
                            double Ginv12 = Ginv[k * t + q];
                            double C12 = Ginv12 + (double)xpx[m] * d[q] * Rinv[k * t + q];
                            rhs0 -= Ginv12 * bb[q];
                            c12b += C12 * bb[q];
                        }
                        double invLhs0 = 1.0 / Ginv11, gHat0 = rhs0 * invLhs0;
                        double invLhs1 = 1.0 / C11;
                        double wr = 0.0;
                        for (int q = 0; q < t; ++q) wr += w[q] * Rinv[q * t + k];
                        double gHat1 = (wr - c12b) * invLhs1;
                        int s0 = 0, s1 = 0;
                        for (int q = 0; q < t; ++q) {
                            int dq = (q == k) ? 0 : (d[q] != 0.0);
                            s0 |= dq << q; s1 |= ((q == k) ? 1 : dq) << q;
                        }
                        double logDelta0 = -0.5 * (log(Ginv11) - gHat0 * gHat0 * Ginv11) + log(bigPi[s0]);
                        double logDelta1 = -0.5 * (log(C11) - gHat1 * gHat1 * C11) + log(bigPi[s1]);
                        double prob1 = 1.0 / (1.0 + exp(logDelta0 - logDelta1));
                        if (uu[k * p + m] < prob1) {                      /* :306 */
                            d[k] = 1.0;
                            float nb = (float)(gHat1 + zz[k * p + m] * sqrt(invLhs1));
                            bb[k] = newa[k] = nb;
                            float dd = (float)olda[k] - nb;
                            for (int64_t q = 0; q < b; ++q) r[k * b + q] += dd * Gm[q * b + jj];   /* :309 */
                        } else {
                            bb[k] = (double)(float)(gHat0 + zz[k * p + m] * sqrt(invLhs0));
                            d[k] = 0.0; newa[k] = 0.0;
                            if (olda[k] != 0.0) {
                                float dd = (float)olda[k];
                                for (int64_t q = 0; q < b; ++q) r[k * b + q] += dd * Gm[q * b + jj]; /* :316 */
                            }
                        }
                    }
                    for (int k = 0; k < t; ++k) {
                        beta[k * p + m] = (float)bb[k]; delta[k * p + m] = (float)d[k]; alpha[k * p + m] = (float)newa[k];
                    }
                }
            }
        for (int k = 0; k < t; ++k)
            for (int64_t a = 0; a < b; ++a) {
                float d = aold[k * b + a] - alpha[k * p + s + a];        /* :329 / :533 */
                if (independent) dsave[k * p + s + a] = d;
                else if (d != 0.0f) jwo_saxpy(d, X + (s + a) * n, ycorr + k * n, n, 1);
            }
        free(Gm); free(r); free(aold);
    }
    if (independent) {                                                   /* :432-436 / :641-645 */
        for (int k = 0; k < t; ++k)
            for (int64_t j = 0; j < p; ++j)
                if (dsave[k * p + j] != 0.0f) jwo_saxpy(dsave[k * p + j], X + j * n, ycorr + k * n, n, 1);
        free(snap); free(dsave);
    }
}

/* ======================================================================== */
/* schedule helpers                                                         */
/* ======================================================================== */
int jwo_bayesr_block_nreps(int64_t iter, int64_t burnin, int64_t block_size) { /* BayesR.jl:22-25 */
    if (block_size < 1) return -1;
    return iter <= burnin ? 1 : (int)block_size;
}
int jwo_validate_block_starts(const int64_t* s, int64_t ns, int64_t nmarkers) { /* JWAS.jl:73-79 */
    if (ns <= 0) return 1;
    if (s[0] != 1) return 2;
    for (int64_t i = 0; i < ns; ++i) if (s[i] < 1 || s[i] > nmarkers) return 3;
    for (int64_t i = 1; i < ns; ++i) if (s[i] <= s[i - 1]) return 4;
    return 0;
}
void jwo_bayesr_sigma_sufficient_statistics(const float* alpha, const int32_t* delta,
                                            const double* gamma, int64_t p, double* ssq, int64_t* nnz) {
    double s = 0.0; int64_t c = 0;                                   /* variance_components.jl:68-79 */
    for (int64_t j = 0; j < p; ++j) {
        int dj = delta[j];
        if (dj > 1) { s += (double)alpha[j] * (double)alpha[j] / gamma[dj - 1]; c += 1; }
    }
    *ssq = s; *nnz = c;
}

/* ======================================================================== */
/* contract-arithmetic hyper-parameter helpers                              */
/* ======================================================================== */

/* BayesB per-marker variance (variance_components.jl:60-66, 169-172):
 *   var_j = (beta_j^2 + df*scale) / chisq(df + 1),  chisq(k) = 2*Gamma(k/2) by Marsaglia-Tsang
 * with draws from the native stream (pseudo-traits 126/127 of the marker's counter space, attempt
 * number in the repetition field).  Twin of jw_k_bayesb_var. */
void jwo_bayesb_variances(const float* beta, int64_t p, double df, double scale,
                          uint64_t seed, uint32_t iter, double* ve) {
    for (int64_t j = 0; j < p; ++j) {
        double shape = 0.5 * (df + 1.0);
        double boost = 1.0;
        if (shape < 1.0) {
            double uu = jw_draw_uniform(seed, (uint32_t)j, iter, 126u, 0u);
            boost = jw_exp(jw_log(uu) / shape);
            shape += 1.0;
        }
        double d = shape - 1.0 / 3.0, c = 1.0 / jw_sqrt(9.0 * d);
        double g = d;
        for (uint32_t att = 0; att < 64u; ++att) {
            double zz = jw_draw_normal(seed, (uint32_t)j, iter, 127u, att);
            double uu = jw_draw_uniform(seed, (uint32_t)j, iter, 127u, att);
            double v = 1.0 + c * zz;
            if (v <= 0.0) continue;
            v = v * v * v;
            if (jw_log(uu) < 0.5 * zz * zz + d - d * v + d * jw_log(v)) { g = d * v; break; }
        }
        double chisq = 2.0 * g * boost;
        double b = (double)beta[j];
        ve[j] = (b * b + df * scale) / chisq;
    }
}

/* ======================================================================== */
/* contract-arithmetic sweep                                                */
/* ======================================================================== */

static double draw_u(const jwo_sweep_args* a, int64_t j, int trait, int rep) {
    if (a->u) return a->u[((int64_t)rep * a->ntraits + trait) * a->p + j];
    return jw_draw_uniform(a->seed, (uint32_t)j, a->iter, (uint32_t)trait, (uint32_t)rep);
}
static double draw_z(const jwo_sweep_args* a, int64_t j, int trait, int rep) {
    if (a->z) return a->z[((int64_t)rep * a->ntraits + trait) * a->p + j];
    return jw_draw_normal(a->seed, (uint32_t)j, a->iter, (uint32_t)trait, (uint32_t)rep);
}

/* one marker of BayesABC.jl:36-55 in binary64; returns Float32 delta-alpha (old-new) */
static float abc_step_contract(double r, float xpx, float* alpha, float* beta, int32_t* delta,
                               double invVarRes, double varEffect, double pi, double u, double z) {
    double x = (double)xpx, aold = (double)(*alpha);
    double rhs = (r + x * aold) * invVarRes;
    double lhs = x * invVarRes + 1.0 / varEffect;
    double invLhs = 1.0 / lhs;
    double gHat = rhs * invLhs;
    double logDelta1 = -0.5 * (jw_log(lhs) + jw_log(varEffect) - gHat * rhs) + jw_log(1.0 - pi);
    double logDelta0 = jw_log(pi);
    float oldA = *alpha, newA;
    if (logDelta0 - logDelta1 < jw_logit_threshold(u)) {      /* rand() < probDelta1, log-odds form */
        *delta = 1;
        newA = (float)(gHat + z * jw_sqrt(invLhs));
        *beta = newA;
    } else {
        *delta = 0;
        *beta = (float)(z * jw_sqrt(varEffect));
        newA = 0.0f;
    }
    *alpha = newA;
    return oldA - newA;
}

/* one marker of BayesR.jl:57-95 in binary64 */
static float r_step_contract(double r, float xpx, float* alpha, int32_t* delta,
                             double invVarRes, double sigmaSq, const double* pi,
                             const double* gamma, int nclasses, double u, double z) {
    double x = (double)xpx, aold = (double)(*alpha);
    double rhs = (r + x * aold) * invVarRes;
    double lp[16], pr[16];
    lp[0] = jw_log(pi[0]);
    for (int k = 1; k < nclasses; ++k) {
        double varEffect = gamma[k] * sigmaSq;
        double lhs = x * invVarRes + 1.0 / varEffect;
        double invLhs = 1.0 / lhs;
        double betaHat = invLhs * rhs;
        lp[k] = 0.5 * (jw_log(invLhs) - jw_log(varEffect) + betaHat * rhs) + jw_log(pi[k]);
    }
    double mx = lp[0];
    for (int k = 1; k < nclasses; ++k) if (lp[k] > mx) mx = lp[k];
    double se = 0.0;
    for (int k = 0; k < nclasses; ++k) { pr[k] = jw_exp(lp[k] - mx); se += pr[k]; }
    /* Categorical(exp(lp - logsumexp)) (BayesR.jl:74-79; Distributions.jl: first i with cumsum > u)
     * drawn on the unnormalised weights e_k = exp(lp_k - max): first class whose cumulative weight
     * exceeds u * sum(e) -- the same event without the second round of exps */
    double target = u * se;
    int cls = 0; double cp = pr[0];
    while (cp <= target && cls < nclasses - 1) { cls += 1; cp += pr[cls]; }
    *delta = cls + 1;
    float oldA = *alpha, newA = 0.0f;
    if (cls > 0) {
        double varEffect = gamma[cls] * sigmaSq;
        double lhs = x * invVarRes + 1.0 / varEffect;
        double invLhs = 1.0 / lhs;
        double betaHat = invLhs * rhs;
        newA = (float)(betaHat + z * jw_sqrt(invLhs));
    }
    *alpha = newA;
    return oldA - newA;
}

/* 2x2 .. 8x8 inverse in binary64 with a FIXED elimination order (no pivoting), so the
 * CUDA side can restate it bit for bit; inputs are SPD covariance matrices. */
static void inv_spd_fixed(const double* A, int t, double* Ai) {
    double M[8][16];
    for (int i = 0; i < t; ++i)
        for (int j = 0; j < t; ++j) { M[i][j] = A[i * t + j]; M[i][t + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < t; ++c) {
        double d = 1.0 / M[c][c];
        for (int j = 0; j < 2 * t; ++j) M[c][j] = M[c][j] * d;
        for (int r = 0; r < t; ++r) if (r != c) {
            double f = M[r][c];
            for (int j = 0; j < 2 * t; ++j) M[r][j] = M[r][j] - f * M[c][j];
        }
    }
    for (int i = 0; i < t; ++i) for (int j = 0; j < t; ++j) Ai[i * t + j] = M[i][t + j];
}

/* ycorr[k] += X_b * d[k] for the markers [s, e) whose stored delta is non-zero, ascending marker order */
static void apply_block_deltas(const jwo_sweep_args* a, const float* dstore, int64_t s, int64_t e,
                               int64_t r0, int64_t r1, float* xbuf) {
    const int64_t n = a->n, p = a->p;
    for (int k = 0; k < a->ntraits; ++k)
        for (int64_t j = s; j < e; ++j) {
            float d = dstore[k * p + j];
            if (d != 0.0f) {
                jwo_decode_marker(a->packed + j * a->stride, n, a->means[j], 1, xbuf);
                float* y = a->ycorr + k * n;
                for (int64_t i = r0; i < r1; ++i) y[i] = fmaf(d, xbuf[i], y[i]);
            }
        }
}

int jwo_sweep_contract(jwo_sweep_args* a) {
    const int64_t n = a->n, p = a->p;
    const int t = a->ntraits;
    const int64_t r0 = a->row_end > 0 ? a->row_begin : 0, r1 = a->row_end > 0 ? a->row_end : n;
    if (t < 1 || t > 8) return 1;
    if (a->method == JWO_METHOD_MT1 && t < 2) return 1;
    if (a->method == JWO_METHOD_MT2 && t != 2) return 1;
    const int lag = a->lag;
    if (lag < 0 || lag > 3) return 1;
    if (lag >= 1 && (a->independent || a->nreps_mode)) return 1;   /* exact schedule only */

    /* fixed-point scale from max|ycorr| over all traits (one S per sweep) */
    float maxabs = 0.0f;
    for (int64_t i = 0; i < n * t; ++i) { float v = fabsf(a->ycorr[i]); if (v > maxabs) maxabs = v; }
    int S = jw_choose_scale_exp(maxabs);
    float scale = jw_pow2f(S);
    double invscale = (double)jw_pow2f(-S);
    a->scale_exp = S;
    a->overflow = 0;

    int32_t* yq = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n * t));
    float* xbuf = (float*)malloc(sizeof(float) * (size_t)n);
    float* dall = (a->independent || lag) ? (float*)calloc((size_t)(p * t), sizeof(float)) : NULL;
    int64_t sq[8];
    double Rinv[64], Ginv[64];
    if (a->method == JWO_METHOD_MT1 || a->method == JWO_METHOD_MT2) {
        inv_spd_fixed(a->Rmat, t, Rinv);
        if (!a->per_marker_G) inv_spd_fixed(a->Gmat, t, Ginv);
    }
    double invVarRes = (a->method == JWO_METHOD_MT1 || a->method == JWO_METHOD_MT2) ? 0.0 : 1.0 / a->vare;

    int have_q = 0;
    for (int64_t ib = 0; ib < a->nblocks; ++ib) {
        int64_t s = a->starts[ib], e = a->starts[ib + 1], b = e - s;
        /* (0) lagged schedule: the updates of block ib-lag-1 reach ycorr only now */
        if (lag && ib >= lag + 1) apply_block_deltas(a, dall, a->starts[ib - lag - 1], a->starts[ib - lag], r0, r1, xbuf);
        /* (1) fixed-point image of ycorr.  Independent blocks all see the entry snapshot
         *     (BayesABC.jl:205, BayesR.jl:209, MTBayesABC.jl:350). */
        if (!a->independent || !have_q) {
            for (int k = 0; k < t; ++k) {
                sq[k] = 0;
                for (int64_t i = 0; i < n; ++i) {
                    int32_t q = (i >= r0 && i < r1) ? jw_quantize(a->ycorr[k * n + i], scale, &a->overflow) : 0;
                    yq[k * n + i] = q; sq[k] += q;
                }
            }
            have_q = 1;
        }
        /* (2) block rhs  r = X_b' ycorr  (tools4genotypes.jl:59-78), exact integers */
        double* r = (double*)malloc(sizeof(double) * (size_t)(b * t));
        float* G = (float*)malloc(sizeof(float) * (size_t)(b * b));
        float* aold = (float*)malloc(sizeof(float) * (size_t)(b * t));
        jwo_gram_block(a->packed, n, a->stride, a->means, s, b, G);
        int64_t* dqv = (int64_t*)calloc((size_t)(2 * b * t + t), sizeof(int64_t));
        int64_t* mqv = dqv + b * t;
        int64_t* sqv = mqv + b * t;
        for (int64_t jj = 0; jj < b; ++jj) {
            const uint8_t* col = a->packed + (s + jj) * a->stride;
            for (int k = 0; k < t; ++k) {
                int64_t dq = 0, mq = 0;
                for (int64_t i = r0; i < r1; ++i) {
                    unsigned c = jw_code(col, i);
                    if (c == 3u) mq += yq[k * n + i]; else dq += (int64_t)c * yq[k * n + i];
                }
                dqv[k * b + jj] = dq; mqv[k * b + jj] = mq;
            }
        }
        for (int k = 0; k < t; ++k) sqv[k] = sq[k];
        if (a->allreduce) a->allreduce(a->ctx, dqv, 2 * b * t + t);     /* C1: per-block rhs all-reduce */
        for (int64_t jj = 0; jj < b; ++jj)
            for (int k = 0; k < t; ++k) {
                double mu = (double)a->means[s + jj];
                r[k * b + jj] = ((double)dqv[k * b + jj] - mu * (double)(sqv[k] - mqv[k * b + jj])) * invscale;
                aold[k * b + jj] = a->alpha[k * p + s + jj];
            }
        free(dqv);
        /* (2b) lagged schedule: block ib-1's updates are not in ycorr yet -- correct the rhs with the
         *      cross-Gram rows of its non-zero deltas, in commit (= marker) order */
        if (lag && ib >= 1) {
            /* blocks ib-lag .. ib-1, oldest first, each in commit (= marker) order */
            for (int64_t ja = a->starts[ib > lag ? ib - lag : 0]; ja < a->starts[ib]; ++ja) {
                int any = 0;
                for (int k = 0; k < t; ++k) any = any || (dall[k * p + ja] != 0.0f);
                if (!any) continue;
                for (int64_t jj = 0; jj < b; ++jj) {
                    jwo_pair q = pair_counts(a->packed + ja * a->stride, a->packed + (s + jj) * a->stride, n);
                    float g = gram_value(q, a->means[ja], a->means[s + jj]);
                    for (int k = 0; k < t; ++k) {
                        float d = dall[k * p + ja];
                        if (d != 0.0f) r[k * b + jj] += (double)d * (double)g;
                    }
                }
            }
        }
        /* (3) in-block chain (BayesABC.jl:153-178, BayesR.jl:146-184, MTBayesABC.jl:276-327) */
        int nreps = a->nreps_mode ? (int)b : 1;
        for (int rep = 0; rep < nreps; ++rep) {
            for (int64_t jj = 0; jj < b; ++jj) {
                int64_t j = s + jj;
                if (a->method == JWO_METHOD_ABC) {
                    float d = abc_step_contract(r[jj], a->xpx[j], &a->alpha[j], &a->beta[j], &a->delta[j],
                                                invVarRes, a->varEffects[j], a->pi[j],
                                                draw_u(a, j, 0, rep), draw_z(a, j, 0, rep));
                    if (d != 0.0f)
                        for (int64_t m = 0; m < b; ++m) r[m] += (double)d * (double)G[jj * b + m];
                } else if (a->method == JWO_METHOD_R) {
                    const double* pij = a->per_marker_pi ? a->pi + j * a->nclasses : a->pi;
                    float d = r_step_contract(r[jj], a->xpx[j], &a->alpha[j], &a->delta[j],
                                              invVarRes, a->sigmaSq, pij, a->gamma, a->nclasses,
                                              draw_u(a, j, 0, rep), draw_z(a, j, 0, rep));
                    if (d != 0.0f)
                        for (int64_t m = 0; m < b; ++m) r[m] += (double)d * (double)G[jj * b + m];
                } else if (a->method == JWO_METHOD_MEGA) {
                    /* megaBayesABC! (BayesABC.jl:1-7): trait k uses vare[k,k], varEffects[k,k], pi[k] */
                    for (int k = 0; k < t; ++k) {
                        float d = abc_step_contract(r[k * b + jj], a->xpx[j], &a->alpha[k * p + j], &a->beta[k * p + j],
                                                    &a->delta[k * p + j], 1.0 / a->Rmat[k * t + k], a->Gmat[k * t + k],
                                                    a->bigPi[k], draw_u(a, j, k, rep), draw_z(a, j, k, rep));
                        if (d != 0.0f)
                            for (int64_t m = 0; m < b; ++m) r[k * b + m] += (double)d * (double)G[jj * b + m];
                    }
                } else if (a->method == JWO_METHOD_MT2) {
                    /* MTBayesABC.jl:163-208 (sampler II, joint states), t = 2, binary64.
                     * States in the order 00,10,01,11 (annotation_setup.jl:18); shared normals z0,z1;
                     * the label is drawn on the unnormalised weights exp(logDelta - max). */
                    double x = (double)a->xpx[j];
                    double w0 = r[jj] + x * (double)a->alpha[j];
                    double w1 = r[b + jj] + x * (double)a->alpha[p + j];
                    double z0 = draw_z(a, j, 0, rep), z1 = draw_z(a, j, 1, rep), uu = draw_u(a, j, 0, rep);
                    double ld[4], bc0[4], bc1[4];
                    for (int st = 0; st < 4; ++st) {
                        double d0 = (double)(st & 1), d1 = (double)((st >> 1) & 1);
                        double l00 = d0 * Rinv[0] * x + Ginv[0];
                        double l01 = (d0 * d1) * Rinv[1] * x + Ginv[1];
                        double l11 = d1 * Rinv[3] * x + Ginv[3];
                        double rhs0 = d0 * (Rinv[0] * w0 + Rinv[2] * w1);
                        double rhs1 = d1 * (Rinv[1] * w0 + Rinv[3] * w1);
                        double det = l00 * l11 - l01 * l01;
                        double i00 = l11 / det, i11 = l00 / det, i01 = -l01 / det;
                        double g0 = i00 * rhs0 + i01 * rhs1, g1 = i01 * rhs0 + i11 * rhs1;
                        ld[st] = -0.5 * (jw_log(det) - (rhs0 * g0 + rhs1 * g1)) + jw_log(a->bigPi[st]);
                        double L00 = jw_sqrt(i00), L10 = i01 / L00, L11 = jw_sqrt(i11 - L10 * L10);
                        bc0[st] = g0 + L00 * z0;
                        bc1[st] = g1 + L10 * z0 + L11 * z1;
                    }
                    double mx = ld[0];
                    for (int st = 1; st < 4; ++st) if (ld[st] > mx) mx = ld[st];
                    double ex[4], se = 0.0;
                    for (int st = 0; st < 4; ++st) { ex[st] = jw_exp(ld[st] - mx); se += ex[st]; }
                    double target = uu * se;
                    int lab = 0; double cp = ex[0];
                    while (cp <= target && lab < 3) { lab += 1; cp += ex[lab]; }
                    float bsel[2] = { (float)bc0[lab], (float)bc1[lab] };
                    for (int k = 0; k < 2; ++k) {
                        int dk = (lab >> k) & 1;
                        float oldA = a->alpha[k * p + j], newA = dk ? bsel[k] : 0.0f;
                        a->alpha[k * p + j] = newA; a->beta[k * p + j] = bsel[k]; a->delta[k * p + j] = dk;
                        float d = oldA - newA;
                        if (d != 0.0f)
                            for (int64_t m = 0; m < b; ++m) r[k * b + m] += (double)d * (double)G[jj * b + m];
                    }
                } else {
                    /* MTBayesABC.jl:78-125, binary64 */
                    double bb[8], olda[8], w[8]; int dd[8];
                    double x = (double)a->xpx[j];
                    if (a->per_marker_G) inv_spd_fixed(a->Gmat + j * t * t, t, Ginv);
                    const double* Pi = a->per_marker_pi ? a->bigPi + j * (1 << t) : a->bigPi;
                    for (int k = 0; k < t; ++k) {
                        bb[k] = (double)a->beta[k * p + j];
                        olda[k] = (double)a->alpha[k * p + j];
                        dd[k] = a->delta[k * p + j] != 0;
                        w[k] = r[k * b + jj] + x * olda[k];
                    }
                    for (int k = 0; k < t; ++k) {
                        double Ginv11 = Ginv[k * t + k];
                        double C11 = Ginv11 + Rinv[k * t + k] * x;
                        double rhs0 = 0.0, c12b = 0.0;
                        for (int q = 0; q < t; ++q) if (q != k) {
                            double Ginv12 = Ginv[k * t + q];
                            double C12 = Ginv12 + x * (double)dd[q] * Rinv[k * t + q];
                            rhs0 = rhs0 - Ginv12 * bb[q];
                            c12b = c12b + C12 * bb[q];
                        }
                        double invLhs0 = 1.0 / Ginv11, gHat0 = rhs0 * invLhs0;
                        double invLhs1 = 1.0 / C11;
                        double wr = 0.0;
                        for (int q = 0; q < t; ++q) wr = wr + w[q] * Rinv[q * t + k];
                        double gHat1 = (wr - c12b) * invLhs1;
                        int s0 = 0, s1 = 0;
                        for (int q = 0; q < t; ++q) {
                            int dq = (q == k) ? 0 : dd[q];
                            s0 |= dq << q; s1 |= ((q == k) ? 1 : dq) << q;
                        }
                        double logDelta0 = -0.5 * (jw_log(Ginv11) - gHat0 * gHat0 * Ginv11) + jw_log(Pi[s0]);
                        double logDelta1 = -0.5 * (jw_log(C11) - gHat1 * gHat1 * C11) + jw_log(Pi[s1]);
                        double uu = draw_u(a, j, k, rep), zz = draw_z(a, j, k, rep);
                        float oldA = a->alpha[k * p + j], newA;
                        if (logDelta0 - logDelta1 < jw_logit_threshold(uu)) {
                            dd[k] = 1;
                            newA = (float)(gHat1 + zz * jw_sqrt(invLhs1));
                            bb[k] = (double)newA;
                        } else {
                            dd[k] = 0;
                            bb[k] = (double)(float)(gHat0 + zz * jw_sqrt(invLhs0));
                            newA = 0.0f;
                        }
                        a->alpha[k * p + j] = newA;
                        a->beta[k * p + j] = (float)bb[k];
                        a->delta[k * p + j] = dd[k];
                        float d = oldA - newA;
                        if (d != 0.0f)
                            for (int64_t m = 0; m < b; ++m)
                                r[k * b + m] += (double)d * (double)G[jj * b + m];
                    }
                }
            }
        }
        /* (4) block exit: ycorr += X_b (alpha_old - alpha)  (BayesABC.jl:181-185), applied
         *     column by column in marker order with one fused multiply-add per element */
        for (int k = 0; k < t; ++k)
            for (int64_t jj = 0; jj < b; ++jj) {
                float d = aold[k * b + jj] - a->alpha[k * p + s + jj];
                if (a->independent || lag) { dall[k * p + s + jj] = d; continue; }
                if (d != 0.0f) {
                    jwo_decode_marker(a->packed + (s + jj) * a->stride, n, a->means[s + jj], 1, xbuf);
                    float* y = a->ycorr + k * n;
                    for (int64_t i = r0; i < r1; ++i) y[i] = fmaf(d, xbuf[i], y[i]);
                }
            }
        free(r); free(G); free(aold);
    }
    if (lag) {                                  /* the updates of the last lag+1 blocks */
        for (int64_t ib = a->nblocks > lag + 1 ? a->nblocks - lag - 1 : 0; ib < a->nblocks; ++ib)
            apply_block_deltas(a, dall, a->starts[ib], a->starts[ib + 1], r0, r1, xbuf);
        free(dall);
    } else if (a->independent) {                /* BayesABC.jl:251-253 */
        for (int k = 0; k < t; ++k)
            for (int64_t j = 0; j < p; ++j) {
                float d = dall[k * p + j];
                if (d != 0.0f) {
                    jwo_decode_marker(a->packed + j * a->stride, n, a->means[j], 1, xbuf);
                    float* y = a->ycorr + k * n;
                    for (int64_t i = r0; i < r1; ++i) y[i] = fmaf(d, xbuf[i], y[i]);
                }
            }
        free(dall);
    }
    free(yq); free(xbuf);
    return a->overflow ? 2 : 0;
}

/* ======================================================================== */
/* contract primitives exposed for tests/test_contract.py                   */
/* ======================================================================== */
double jwo_c_log(double x) { return jw_log(x); }
double jwo_c_exp(double x) { return jw_exp(x); }
double jwo_c_cos2pi(double v) { return jw_cos2pi(v); }
double jwo_c_normal(double u1, double u2) { return jw_normal(u1, u2); }
void jwo_c_philox(const uint32_t* c, const uint32_t* k, uint32_t* out) {
    jw_u32x4 r = jw_philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}
void jwo_c_draws(uint64_t seed, uint32_t iter, uint32_t trait, uint32_t rep, int64_t p, double* u, double* z) {
    for (int64_t j = 0; j < p; ++j) {
        u[j] = jw_draw_uniform(seed, (uint32_t)j, iter, trait, rep);
        z[j] = jw_draw_normal(seed, (uint32_t)j, iter, trait, rep);
    }
}
int32_t jwo_c_quantize(float y, float scale, int* ovf) { return jw_quantize(y, scale, ovf); }
int jwo_c_scale_exp(float maxabs) { return jw_choose_scale_exp(maxabs); }
