// jw_sweep_kernels.cuh -- multi-kernel sweep engine ("engine 0"): one launch per phase.
//   quantise ycorr -> block rhs (packed GEMV, exact int64) -> in-block Gibbs chain on the Gram
//   block (speculative parallel rounds) -> block exit ycorr += X_b * d_alpha.
// Replaces bayesabc_update_marker!/BayesABC!/BayesABC_block!/_independent! (BayesABC.jl:24-255),
// BayesR!/BayesR_block!/_independent! (BayesR.jl:45-273), _MTBayesABC_samplerI! + block +
// independent (MTBayesABC.jl:57-127, 243-437), block_rhs! (tools4genotypes.jl:59-78).
// Arithmetic contract: include/jwas_contract.h; bit-for-bit twin: oracle/jwas_oracle.c
// (jwo_sweep_contract).
#pragma once
#include "jw_common.cuh"

// ------------------------------------------------------------------------------------------
// ycorr -> fixed point.  yq = rint(y * 2^S); sq[k] = sum_i yq (exact, atomics commute).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jw_k_quantize(const float* __restrict__ y, int64_t n, int t, float scale,
              int32_t* __restrict__ yq, long long* __restrict__ sq, int32_t* __restrict__ flags,
              int64_t r0, int64_t r1) {
    int k = blockIdx.y;
    long long acc = 0;
    int ovf = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // rows owned by other ranks contribute nothing to this rank's partial sums
        int32_t q = (i >= r0 && i < r1) ? jw_quantize(y[k * n + i], scale, &ovf) : 0;
        yq[k * n + i] = q;
        acc += q;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd((unsigned long long*)&sq[k], (unsigned long long)acc);
    if (ovf) atomicOr(&flags[0], 1);
}

// ------------------------------------------------------------------------------------------
// block rhs: dq[k][j] = sum_i value(code_ij) * yq[k][i], mq[k][j] = sum over missing calls.
// One warp per (marker, row slab); each lane decodes one 32-bit word (16 individuals).
// ------------------------------------------------------------------------------------------
#define JW_DOT_SLAB_WORDS 256   // 4096 individuals per slab
template <int T, bool MISSING>
__global__ void __launch_bounds__(256)
jw_k_block_dot(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t n, int64_t p,
               int64_t j0, int64_t nj, const int32_t* __restrict__ yq,
               long long* __restrict__ dq, long long* __restrict__ mq, int64_t w0, int64_t w1) {
    int64_t jj = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (jj >= nj) return;
    int64_t j = j0 + jj;
    int lane = threadIdx.x & 31;
    const uint32_t* col = reinterpret_cast<const uint32_t*>(packed + j * stride_d);
    int64_t wbeg = w0 + (int64_t)blockIdx.y * JW_DOT_SLAB_WORDS;
    int64_t wend = min(wbeg + JW_DOT_SLAB_WORDS, w1);
    long long acc[T], macc[T];
#pragma unroll
    for (int k = 0; k < T; ++k) { acc[k] = 0; macc[k] = 0; }
    for (int64_t w = wbeg + lane; w < wend; w += 32) {
        uint32_t v = __ldg(col + w);
        int64_t i0 = w << 4;
        if (i0 + 16 <= n) {
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const int4* yp = reinterpret_cast<const int4*>(yq + k * n + i0);
                bool al = ((k * n + i0) & 3) == 0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    int4 q;
                    if (al) q = __ldg(yp + g);
                    else { const int32_t* s = yq + k * n + i0 + 4 * g; q = make_int4(s[0], s[1], s[2], s[3]); }
                    uint32_t c0 = (v >> (8 * g)) & 3u, c1 = (v >> (8 * g + 2)) & 3u,
                             c2 = (v >> (8 * g + 4)) & 3u, c3 = (v >> (8 * g + 6)) & 3u;
                    if (MISSING) {
                        long long m = 0;
                        if (c0 == 3u) { m += q.x; c0 = 0; }
                        if (c1 == 3u) { m += q.y; c1 = 0; }
                        if (c2 == 3u) { m += q.z; c2 = 0; }
                        if (c3 == 3u) { m += q.w; c3 = 0; }
                        macc[k] += m;
                    }
                    // 4 products of magnitude <= 2^27 fit an int; widen once per group
                    long long s4 = (long long)((int)c0 * q.x) + (long long)((int)c1 * q.y)
                                 + (long long)((int)c2 * q.z) + (long long)((int)c3 * q.w);
                    acc[k] += s4;
                }
            }
        } else {
            for (int e = 0; e < 16 && i0 + e < n; ++e) {
                uint32_t c = (v >> (2 * e)) & 3u;
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    int q = yq[k * n + i0 + e];
                    if (c == 3u) macc[k] += q; else acc[k] += (long long)((int)c * q);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < T; ++k) {
        long long a = acc[k], m = macc[k];
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            if (MISSING) m += __shfl_xor_sync(0xffffffffu, m, o);
        }
        if (lane == 0) {
            if (a != 0) atomicAdd((unsigned long long*)&dq[k * p + j], (unsigned long long)a);
            if (MISSING && m != 0) atomicAdd((unsigned long long*)&mq[k * p + j], (unsigned long long)m);
        }
    }
}

// ------------------------------------------------------------------------------------------
// In-block Gibbs chain.  One CTA per block, one thread per marker.  The sequential chain of the
// reference is executed as speculative rounds: every uncommitted marker evaluates its draw with
// the current rhs; markers that neither hold nor acquire an effect do not touch the rhs
// (BayesABC.jl:50-52), so everything before the first "active" marker commits at once, the
// active marker commits, its Gram row is applied, and the rest re-evaluate.  Draws are indexed
// by (marker, iter, trait, rep), so the result is exactly the sequential one.
// ------------------------------------------------------------------------------------------
struct jw_chain_args {
    int64_t n, p;
    int t, method, nreps_mode, nclasses;
    int per_marker_pi, per_marker_G;
    int block0;                    // first block handled by blockIdx.x == 0
    const int64_t* starts;
    const int64_t* gram_off;
    const float* gram;
    const float* means;
    const float* xpx;
    const long long* dq; const long long* mq; const long long* sq;
    double invscale;
    float* alpha; float* beta; int32_t* delta; float* dalpha;
    double vare, sigmaSq;
    const double* ve; const double* pi;
    double gamma[JW_MAX_CLASSES];
    double Rinv[16]; double Ginv[16];
    const double* Gmat;            // per-marker covariance (p*t*t) when per_marker_G
    const double* bigPi;
    uint64_t seed; uint32_t iter;
    const double* u; const double* z;
    // BayesABC draw-independent terms precomputed for repetition 0 by jw_k_prep_abc (all SMs)
    // instead of inside the one chain CTA: [0]log(1/u-1) [1]z*sqrt(invLhs) [2]invLhs [3]log(lhs)+log(ve)
    // [4]log(1-pi) [5]log(pi), each p doubles; prep_beta0 = float(z*sqrt(ve)).  NULL = inline.
    const double* prep; const float* prep_beta0;
    const double* draws_u; const double* draws_z;   // jw_k_prep_draws output (t*p each) or NULL
    // BayesR / multi-trait: rhs-independent terms.  prep_rm (jw_k_prep_rm): BayesR [(c-1)*p+j] = invLhs_c,
    // [(K-1+c-1)*p+j] = log(invLhs_c)-log(varEffect_c); multi-trait [k*p+j] = log(C11_k).  The logs of the
    // global priors are evaluated once on the host with the same jw_log.
    const double* prep_rm; int host_logs;
    double lpi[JW_MAX_CLASSES]; double mt_lG[JW_MAX_TRAITS]; double mt_lPi[16];
    // lagged schedule: the previous block's updates are not in ycorr yet; its ordered active list and
    // the cross-Gram X_{k-1}'X_k (rows = previous block's markers) correct the rhs of this block
    const float* xgram; const int32_t* xlist; const int32_t* xcount; int64_t xstart;
    const float* xgram_next; int b_next;    // cross-Gram towards the next block: rows are prefetched at commit
    int32_t* act_idx; int32_t* act_cnt;     // ordered active list of this launch (single-block mode)
    int write_active_list;
    unsigned long long* counters;
    int timers;                    // 1 = accumulate phase timers (tools/phase_probe.py); off in production
};

__device__ __forceinline__ double jw_get_u(const jw_chain_args& A, int64_t j, int trait, int rep) {
    if (A.u) return A.u[((int64_t)rep * A.t + trait) * A.p + j];
    return jw_draw_uniform(A.seed, (uint32_t)j, A.iter, (uint32_t)trait, (uint32_t)rep);
}
__device__ __forceinline__ double jw_get_z(const jw_chain_args& A, int64_t j, int trait, int rep) {
    if (A.z) return A.z[((int64_t)rep * A.t + trait) * A.p + j];
    return jw_draw_normal(A.seed, (uint32_t)j, A.iter, (uint32_t)trait, (uint32_t)rep);
}

// fixed-order Gauss-Jordan, twin of inv_spd_fixed in the oracle
__host__ __device__ inline void jw_inv_spd_fixed(const double* A, int t, double* Ai) {
    double M[JW_MAX_TRAITS][2 * JW_MAX_TRAITS];
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

__device__ __forceinline__ int jw_categorical(const double* probs, int k, double u) {
    double cp = probs[0]; int i = 0;
    while (cp <= u && i < k - 1) { i += 1; cp += probs[i]; }
    return i;
}

// draws of repetition 0 for every (trait, marker), computed by the whole GPU instead of inside the one
// chain CTA: du = log-odds threshold of the uniform (BayesR: the uniform itself), dz = standard normal
__global__ void __launch_bounds__(256)
jw_k_prep_draws(jw_chain_args A, double* __restrict__ du, double* __restrict__ dz) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.p * A.t) return;
    const int k = (int)(idx / A.p); const int64_t j = idx % A.p;
    double u = jw_get_u(A, j, k, 0);
    du[idx] = (A.method == 1 || A.method == 3) ? u : jw_logit_threshold(u);
    dz[idx] = jw_get_z(A, j, k, 0);
}

__global__ void __launch_bounds__(256)
jw_k_prep_rm(jw_chain_args A, double* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = A.p;
    if (j >= p) return;
    const double x = (double)A.xpx[j];
    if (A.method == 1) {
        const double invVarRes = 1.0 / A.vare;
        const int K1 = A.nclasses - 1;
        for (int c = 1; c < A.nclasses; ++c) {
            double varEffect = A.gamma[c] * A.sigmaSq;
            double lhs = x * invVarRes + 1.0 / varEffect;
            double invLhs = 1.0 / lhs;
            out[(int64_t)(c - 1) * p + j] = invLhs;
            out[(int64_t)(K1 + c - 1) * p + j] = jw_log(invLhs) - jw_log(varEffect);
        }
    } else {
        for (int k = 0; k < A.t; ++k)
            out[(int64_t)k * p + j] = jw_log(A.Ginv[k * A.t + k] + A.Rinv[k * A.t + k] * x);
    }
}

__global__ void __launch_bounds__(256)
jw_k_prep_abc(jw_chain_args A, double* __restrict__ prep, float* __restrict__ beta0) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = A.p;
    if (j >= p) return;
    const double x = (double)A.xpx[j];
    const double invVarRes = 1.0 / A.vare;
    const double ve = A.ve[j], pi = A.pi[j];
    const double lhs = x * invVarRes + 1.0 / ve;
    const double invLhs = 1.0 / lhs;
    const double u = jw_get_u(A, j, 0, 0), z = jw_get_z(A, j, 0, 0);
    prep[j] = jw_logit_threshold(u);
    prep[p + j] = z * jw_sqrt(invLhs);
    prep[2 * p + j] = invLhs;
    prep[3 * p + j] = jw_log(lhs) + jw_log(ve);
    prep[4 * p + j] = jw_log(1.0 - pi);
    prep[5 * p + j] = jw_log(pi);
    beta0[j] = (float)(z * jw_sqrt(ve));
}

// ------------------------------------------------------------------------------------------
// One marker's single-site conditional, shared by every chain driver (jw_chain_block and the
// pipelined jw_chain_unit): load_constants() gathers everything that does not depend on the rhs,
// load_draws() the repetition's uniform / normal, eval() turns (rhs, current state) into the new
// state and says whether the marker's effect changes (the only case that touches the rhs).
//   METHOD 0 BayesABC (BayesABC.jl:24-58)      1 BayesR (BayesR.jl:58-95)
//          2 multi-trait sampler I (MTBayesABC.jl:78-125)   3 sampler II (MTBayesABC.jl:163-208)
//          4 megaBayesABC (BayesABC.jl:1-7)
// ------------------------------------------------------------------------------------------
template <int METHOD, int T>
struct jw_marker_eval {
    double x, invVarRes;
    double Ginv[T * T];
    // METHOD 0: draw-independent constants (from jw_k_prep_abc when available)
    double c_invLhs, c_L, c_lpc, c_lp0, c_ve;
    bool use_prep;
    double u0, zs1_0; float beta0_0;
    // METHOD 1: class terms that do not depend on the rhs (BayesR.jl:64-72); exactly JW_R_CLASSES = 4
    // mixture classes (BAYESR_GAMMA, JWAS.jl:12), loops fully unrolled so the terms live in registers
    double rc_invLhs[JW_R_CLASSES], rc_base[JW_R_CLASSES], rc_lpi[JW_R_CLASSES];
    // METHOD 2: logs of the per-marker constants (MTBayesABC.jl:88-105)
    double mt_lG[T], mt_lC[T], mt_lPi[1 << T];
    // draws of the current repetition (u[] holds the log-odds threshold except for METHOD 1 / 3)
    double u[T], z[T];
    double zs1; float beta0;                   // METHOD 0: z*sqrt(invLhs), float(z*sqrt(ve))

    __device__ __forceinline__ void load_constants(const jw_chain_args& A, const int64_t j) {
        const int64_t p = A.p;
        x = (double)A.xpx[j];
        if (METHOD == 2 || METHOD == 3) {
            if (A.per_marker_G) jw_inv_spd_fixed(A.Gmat + j * T * T, T, Ginv);
            else for (int q = 0; q < T * T; ++q) Ginv[q] = A.Ginv[q];
        }
        invVarRes = (METHOD == 2 || METHOD == 3 || METHOD == 4) ? 0.0 : 1.0 / A.vare;
        c_invLhs = 0; c_L = 0; c_lpc = 0; c_lp0 = 0; c_ve = 1;
        use_prep = (METHOD == 0) && (A.prep != nullptr);
        u0 = 0.0; zs1_0 = 0.0; beta0_0 = 0.0f;
        if (METHOD == 0) {
            if (use_prep) {
                c_invLhs = A.prep[2 * p + j]; c_L = A.prep[3 * p + j];
                c_lpc = A.prep[4 * p + j]; c_lp0 = A.prep[5 * p + j];
                u0 = A.prep[j]; zs1_0 = A.prep[p + j]; beta0_0 = A.prep_beta0[j];
            } else {
                c_ve = A.ve[j];
                double pi = A.pi[j];
                double c_lhs = x * invVarRes + 1.0 / c_ve;
                c_invLhs = 1.0 / c_lhs;
                c_L = jw_log(c_lhs) + jw_log(c_ve);
                c_lpc = jw_log(1.0 - pi);
                c_lp0 = jw_log(pi);
            }
        }
        if (METHOD == 1) {
            rc_invLhs[0] = 0.0; rc_base[0] = 0.0;
            if (A.prep_rm != nullptr && A.host_logs) {
                const int K1 = JW_R_CLASSES - 1;
                rc_lpi[0] = A.lpi[0];
#pragma unroll
                for (int c = 1; c < JW_R_CLASSES; ++c) {
                    rc_invLhs[c] = A.prep_rm[(int64_t)(c - 1) * p + j];
                    rc_base[c] = A.prep_rm[(int64_t)(K1 + c - 1) * p + j];
                    rc_lpi[c] = A.lpi[c];
                }
            } else {
                const double* pij = A.per_marker_pi ? A.pi + j * JW_R_CLASSES : A.pi;
                rc_lpi[0] = jw_log(pij[0]);
#pragma unroll
                for (int c = 1; c < JW_R_CLASSES; ++c) {
                    double varEffect = A.gamma[c] * A.sigmaSq;
                    double lhs = x * invVarRes + 1.0 / varEffect;
                    rc_invLhs[c] = 1.0 / lhs;
                    rc_base[c] = jw_log(rc_invLhs[c]) - jw_log(varEffect);
                    rc_lpi[c] = jw_log(pij[c]);
                }
            }
        }
        if (METHOD == 2) {
            if (A.prep_rm != nullptr && A.host_logs) {
#pragma unroll
                for (int k = 0; k < T; ++k) { mt_lG[k] = A.mt_lG[k]; mt_lC[k] = A.prep_rm[(int64_t)k * p + j]; }
#pragma unroll
                for (int q = 0; q < (1 << T); ++q) mt_lPi[q] = A.mt_lPi[q];
            } else {
                const double* Pi = A.per_marker_pi ? A.bigPi + j * (1 << T) : A.bigPi;
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    mt_lG[k] = jw_log(Ginv[k * T + k]);
                    mt_lC[k] = jw_log(Ginv[k * T + k] + A.Rinv[k * T + k] * x);
                }
#pragma unroll
                for (int q = 0; q < (1 << T); ++q) mt_lPi[q] = jw_log(Pi[q]);
            }
        }
    }

    __device__ __forceinline__ void load_draws(const jw_chain_args& A, const int64_t j, const int rep) {
        const int64_t p = A.p;
        zs1 = 0.0; beta0 = 0.0f;
        if (METHOD == 0 && use_prep && rep == 0) {
            u[0] = u0; zs1 = zs1_0; beta0 = beta0_0; z[0] = 0.0;
        } else {
#pragma unroll
            for (int k = 0; k < T; ++k) {
                if (rep == 0 && A.draws_u != nullptr) {
                    u[k] = A.draws_u[k * p + j]; z[k] = A.draws_z[k * p + j];
                } else {
                    u[k] = jw_get_u(A, j, k, rep); z[k] = jw_get_z(A, j, k, rep);
                    if (METHOD != 1 && METHOD != 3) u[k] = jw_logit_threshold(u[k]);
                }
            }
            if (METHOD == 0) {
                if (use_prep) c_ve = A.ve[j];
                zs1 = z[0] * jw_sqrt(c_invLhs); beta0 = (float)(z[0] * jw_sqrt(c_ve));
            }
        }
    }

    // returns true when the marker's effect changes (a_cur != newA for some trait)
    __device__ __forceinline__ bool eval(const jw_chain_args& A, const double* r, const float* a_cur,
                                         const float* b_cur, const int* d_cur,
                                         float* newA, float* newB, int* newD) const {
        bool active = false;
        if (METHOD == 0) {
            double aold = (double)a_cur[0];
            double rhs = (r[0] + x * aold) * invVarRes;
            double gHat = rhs * c_invLhs;
            double logDelta1 = -0.5 * (c_L - gHat * rhs) + c_lpc;
            if (c_lp0 - logDelta1 < u[0]) {           // u[] holds the log-odds threshold of the draw
                newD[0] = 1;
                newA[0] = (float)(gHat + zs1);
                newB[0] = newA[0];
            } else {
                newD[0] = 0;
                newB[0] = beta0;
                newA[0] = 0.0f;
            }
            active = (a_cur[0] - newA[0]) != 0.0f;
        } else if (METHOD == 1) {
            double aold = (double)a_cur[0];
            double rhs = (r[0] + x * aold) * invVarRes;
            double lp[JW_R_CLASSES], ex[JW_R_CLASSES];
            lp[0] = rc_lpi[0];
#pragma unroll
            for (int c = 1; c < JW_R_CLASSES; ++c) {
                double betaHat = rc_invLhs[c] * rhs;
                lp[c] = 0.5 * (rc_base[c] + betaHat * rhs) + rc_lpi[c];
            }
            double mx = lp[0];
#pragma unroll
            for (int c = 1; c < JW_R_CLASSES; ++c) if (lp[c] > mx) mx = lp[c];
            double se = 0.0;
#pragma unroll
            for (int c = 0; c < JW_R_CLASSES; ++c) { ex[c] = jw_exp(lp[c] - mx); se += ex[c]; }
            // Categorical(exp(lp - logsumexp)) (BayesR.jl:74-79) drawn on the unnormalised weights:
            // first class whose cumulative weight exceeds u * sum
            const double target = u[0] * se;
            int cls = 0; double cp = ex[0];
#pragma unroll
            for (int c = 1; c < JW_R_CLASSES; ++c) {
                if (cls == c - 1 && cp <= target) { cls = c; cp += ex[c]; }
            }
            newD[0] = cls + 1;
            newA[0] = 0.0f;
            if (cls > 0) {
                const double il = cls == 1 ? rc_invLhs[1] : (cls == 2 ? rc_invLhs[2] : rc_invLhs[3]);
                double betaHat = il * rhs;
                newA[0] = (float)(betaHat + z[0] * jw_sqrt(il));
            }
            newB[0] = 0.0f;
            active = (a_cur[0] - newA[0]) != 0.0f;
        } else if (METHOD == 4) {
            // megaBayesABC! (BayesABC.jl:1-7): T independent BayesABC steps; trait k uses
            // vare = 1/Rinv[k] (diagonal stored), varEffect = Ginv[k] (variance itself), pi = lpi[k]
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const double ivr = A.Rinv[k], vek = A.Ginv[k], pik = A.lpi[k];
                const double aold = (double)a_cur[k];
                const double rhs = (r[k] + x * aold) * ivr;
                const double lhs = x * ivr + 1.0 / vek;
                const double invLhs = 1.0 / lhs;
                const double gHat = rhs * invLhs;
                const double logDelta1 = -0.5 * (jw_log(lhs) + jw_log(vek) - gHat * rhs) + jw_log(1.0 - pik);
                const double logDelta0 = jw_log(pik);
                if (logDelta0 - logDelta1 < u[k]) {
                    newD[k] = 1;
                    newA[k] = (float)(gHat + z[k] * jw_sqrt(invLhs));
                    newB[k] = newA[k];
                } else {
                    newD[k] = 0;
                    newB[k] = (float)(z[k] * jw_sqrt(vek));
                    newA[k] = 0.0f;
                }
                if ((a_cur[k] - newA[k]) != 0.0f) active = true;
            }
        } else if (METHOD == 3) {
            // MTBayesABC.jl:163-208 (sampler II, joint states), T == 2; twin of the oracle's MT2 step
            const double w0 = r[0] + x * (double)a_cur[0];
            const double w1 = r[T - 1] + x * (double)a_cur[T - 1];
            const double z0 = z[0], z1 = z[T - 1], uu = u[0];
            double ld[4], bc0[4], bc1[4];
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                const double d0 = (double)(st & 1), d1 = (double)((st >> 1) & 1);
                const double l00 = d0 * A.Rinv[0] * x + Ginv[0];
                const double l01 = (d0 * d1) * A.Rinv[1] * x + Ginv[1];
                const double l11 = d1 * A.Rinv[3] * x + Ginv[T * T - 1];
                const double rhs0 = d0 * (A.Rinv[0] * w0 + A.Rinv[2] * w1);
                const double rhs1 = d1 * (A.Rinv[1] * w0 + A.Rinv[3] * w1);
                const double det = l00 * l11 - l01 * l01;
                const double i00 = l11 / det, i11 = l00 / det, i01 = -l01 / det;
                const double g0 = i00 * rhs0 + i01 * rhs1, g1 = i01 * rhs0 + i11 * rhs1;
                ld[st] = -0.5 * (jw_log(det) - (rhs0 * g0 + rhs1 * g1)) + jw_log(A.bigPi[st]);
                const double L00 = jw_sqrt(i00), L10 = i01 / L00, L11 = jw_sqrt(i11 - L10 * L10);
                bc0[st] = g0 + L00 * z0;
                bc1[st] = g1 + L10 * z0 + L11 * z1;
            }
            double mx = ld[0];
#pragma unroll
            for (int st = 1; st < 4; ++st) if (ld[st] > mx) mx = ld[st];
            double ex[4], se = 0.0;
#pragma unroll
            for (int st = 0; st < 4; ++st) { ex[st] = jw_exp(ld[st] - mx); se += ex[st]; }
            const double target = uu * se;
            int lab = 0; double cp = ex[0];
#pragma unroll
            for (int c = 1; c < 4; ++c) { if (lab == c - 1 && cp <= target) { lab = c; cp += ex[c]; } }
            const float b0 = (float)(lab == 0 ? bc0[0] : lab == 1 ? bc0[1] : lab == 2 ? bc0[2] : bc0[3]);
            const float b1 = (float)(lab == 0 ? bc1[0] : lab == 1 ? bc1[1] : lab == 2 ? bc1[2] : bc1[3]);
            newB[0] = b0; newB[T - 1] = b1;
            newD[0] = lab & 1; newD[T - 1] = (lab >> 1) & 1;
            newA[0] = newD[0] ? b0 : 0.0f; newA[T - 1] = newD[T - 1] ? b1 : 0.0f;
            active = ((a_cur[0] - newA[0]) != 0.0f) || ((a_cur[T - 1] - newA[T - 1]) != 0.0f);
        } else {
            // MTBayesABC.jl:78-125
            double bb[T], olda[T], w[T]; int dd[T];
#pragma unroll
            for (int k = 0; k < T; ++k) {
                bb[k] = (double)b_cur[k]; olda[k] = (double)a_cur[k]; dd[k] = d_cur[k] != 0;
                w[k] = r[k] + x * olda[k];
            }
#pragma unroll
            for (int k = 0; k < T; ++k) {
                double Ginv11 = Ginv[k * T + k];
                double C11 = Ginv11 + A.Rinv[k * T + k] * x;
                double rhs0 = 0.0, c12b = 0.0;
#pragma unroll
                for (int q = 0; q < T; ++q) if (q != k) {
                    double Ginv12 = Ginv[k * T + q];
                    double C12 = Ginv12 + x * (double)dd[q] * A.Rinv[k * T + q];
                    rhs0 = rhs0 - Ginv12 * bb[q];
                    c12b = c12b + C12 * bb[q];
                }
                double invLhs0 = 1.0 / Ginv11, gHat0 = rhs0 * invLhs0;
                double invLhs1 = 1.0 / C11;
                double wr = 0.0;
#pragma unroll
                for (int q = 0; q < T; ++q) wr = wr + w[q] * A.Rinv[q * T + k];
                double gHat1 = (wr - c12b) * invLhs1;
                int s0 = 0, s1 = 0;
#pragma unroll
                for (int q = 0; q < T; ++q) {
                    int dqq = (q == k) ? 0 : dd[q];
                    s0 |= dqq << q; s1 |= ((q == k) ? 1 : dqq) << q;
                }
                double logDelta0 = -0.5 * (mt_lG[k] - gHat0 * gHat0 * Ginv11) + mt_lPi[s0];
                double logDelta1 = -0.5 * (mt_lC[k] - gHat1 * gHat1 * C11) + mt_lPi[s1];
                if (logDelta0 - logDelta1 < u[k]) {
                    dd[k] = 1;
                    newA[k] = (float)(gHat1 + z[k] * jw_sqrt(invLhs1));
                    bb[k] = (double)newA[k];
                } else {
                    dd[k] = 0;
                    bb[k] = (double)(float)(gHat0 + z[k] * jw_sqrt(invLhs0));
                    newA[k] = 0.0f;
                }
                newB[k] = (float)bb[k]; newD[k] = dd[k];
                if ((a_cur[k] - newA[k]) != 0.0f) active = true;
            }
        }
        return active;
    }
};

struct jw_no_wait { __device__ __forceinline__ bool operator()() const { return true; } };

// what changes from block to block (kept small so the big jw_chain_args can stay in parameter space)
struct jw_chain_blk {
    const long long* sq;
    int32_t* act_idx; int32_t* act_cnt; int write_active_list;
    const float* xgram; const int32_t* xlist; const int32_t* xcount; int64_t xstart;
    const float* xgram2;                    // lag 2 (pipelined chain): cross-Gram X_{k-2}'X_k, NULL otherwise
    const float* xgram_next; int b_next;
    int64_t prefetch_s; int prefetch_b;     // next block's chain inputs to pull towards L2 (0 = none)
    int64_t s; int b; int64_t gram_off;     // this block: first marker, size, Gram offset (b = 0: look them up)
    // (c) previous block's commits kept in shared memory by the dedicated chain CTA (lagged schedule)
    int xcount_smem;                        // >= 0: number of entries in the shared list; -1: use xlist/xcount
    // multi-GPU fused sweep: the block's partial rhs of every rank, pushed over NVLink into this GPU's
    // exchange slots: [rank][ dq: T*slot_b | mq: T*slot_b | sq: T ] 16-byte self-validating words
    // {lo32, tag, hi32, tag}; NULL = single GPU
    const uint4* xslots; int xworld; int64_t slot_stride; int slot_b; unsigned xtag; int32_t* xflags;
};

// ---- 16-byte words of the multi-GPU exchange: value + tag in each 8-byte half, so a word is complete in
//      itself -- one vector store over NVLink, no flag, no fence; readers poll until both tags match ----
__device__ __forceinline__ void jw_ll_store(void* dst, long long v, unsigned tag) {
    const unsigned lo = (unsigned)(unsigned long long)v, hi = (unsigned)((unsigned long long)v >> 32);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(dst), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 jw_ll_load(const void* src) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
    return v;
}
// exact int64 sum over the W ranks' copies of one value (copies `stride` words apart): the loads of up to JW_LL_CHUNK
// ranks are issued together, ranks that have not arrived are polled again.  false = the sweep was abandoned.
#ifndef JW_LL_CHUNK
#define JW_LL_CHUNK 4
#endif
#ifndef JW_LL_INLINE
#define JW_LL_INLINE __forceinline__
#endif
__device__ JW_LL_INLINE bool jw_ll_sum(const uint4* base, const int64_t stride, const int W, const unsigned tag,
                                       int32_t* flags, long long& out) {
    long long acc = 0;
    unsigned spins = 0; unsigned long long t0 = 0;
    for (int r0 = 0; r0 < W; r0 += JW_LL_CHUNK) {
        const int nr = min(JW_LL_CHUNK, W - r0);
        unsigned pend = (1u << nr) - 1u;
        while (pend) {
            uint4 v[JW_LL_CHUNK];
#pragma unroll
            for (int q = 0; q < JW_LL_CHUNK; ++q) if ((pend >> q) & 1u) v[q] = jw_ll_load(base + (int64_t)(r0 + q) * stride);
#pragma unroll
            for (int q = 0; q < JW_LL_CHUNK; ++q)
                if (((pend >> q) & 1u) && v[q].y == tag && v[q].w == tag) {
                    acc += (long long)(((unsigned long long)v[q].z << 32) | (unsigned long long)v[q].x);
                    pend &= ~(1u << q);
                }
            if (pend && (++spins & 255u) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                int ab;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(ab) : "l"(flags + 2) : "memory");
                if (ab != 0) return false;
                if (now - t0 > 20000000000ull) { atomicExch(&flags[2], 1); return false; }
            }
        }
    }
    out = acc;
    return true;
}
// the three integer sums a marker's rhs needs, over all ranks
__device__ JW_LL_INLINE bool jw_ll_rhs(const jw_chain_blk& B, const int T, const int k, const int m, const bool has_mq,
                                          long long& dq, long long& mq, long long& sqk) {
    bool ok = jw_ll_sum(B.xslots + (int64_t)k * B.slot_b + m, B.slot_stride, B.xworld, B.xtag, B.xflags, dq);
    mq = 0;
    if (has_mq) ok = jw_ll_sum(B.xslots + (int64_t)(T + k) * B.slot_b + m, B.slot_stride, B.xworld, B.xtag, B.xflags, mq) && ok;
    ok = jw_ll_sum(B.xslots + (int64_t)2 * T * B.slot_b + k, B.slot_stride, B.xworld, B.xtag, B.xflags, sqk) && ok;
    return ok;
}
__device__ __forceinline__ jw_chain_blk jw_chain_blk_from(const jw_chain_args& A) {
    jw_chain_blk B;
    B.sq = A.sq; B.act_idx = A.act_idx; B.act_cnt = A.act_cnt; B.write_active_list = A.write_active_list;
    B.xgram = nullptr; B.xgram2 = nullptr; B.xlist = nullptr; B.xcount = nullptr; B.xstart = 0; B.xgram_next = nullptr; B.b_next = 0;
    B.prefetch_s = 0; B.prefetch_b = 0; B.s = 0; B.b = 0; B.gram_off = 0; B.xcount_smem = -1;
    B.xslots = nullptr; B.xworld = 1; B.slot_stride = 0; B.slot_b = 0; B.xtag = 0; B.xflags = nullptr;
    return B;
}

// wait_fn() is called after everything that does not depend on the block rhs has been loaded
// (state, constants, Gram-row prefetches): the fused engine spins there for the other CTAs'
// partial sums, so those global-memory latencies hide behind the wait.  Returns false on abort.
// Shared memory of the chain (dynamic, carved from a caller-supplied base):
//   wmin[2][32] | cnt[33] | dc[T][1024] | list_idx[cap] | list_d[T][cap]
// The commit list exists only for panels larger than one thread-block of markers (cap = panel size).
#define JW_CHAIN_SB 1024
__host__ __device__ inline size_t jw_chain_smem_bytes(int T, int list_cap, int nlists = 1) {
    return 256 + 256 + (size_t)T * JW_CHAIN_SB * 4 + (size_t)nlists * list_cap * 4 * (1 + T);
}

// wait_fn() is called after everything that does not depend on the block rhs has been loaded
// (state, constants, Gram-row prefetches): the fused engine spins there for the other CTAs'
// partial sums, so those global-memory latencies hide behind the wait.  Returns false on abort.
//
// Panels larger than the thread block are walked in sub-blocks of blockDim.x markers; every commit is
// appended to a list so that a later sub-block starts from  base rhs + sum_commits d*G[commit][j]
// (added in commit order: the same sums, in the same order, as the one-thread-per-marker chain).
template <int METHOD, int T, bool MULTI, class WaitFn>
__device__ __forceinline__ int jw_chain_block(const jw_chain_args& A, const jw_chain_blk& B, const int ib,
                                              WaitFn wait_fn, unsigned char* smem_base, const int list_cap) {
    int (*s_wmin)[32] = reinterpret_cast<int (*)[32]>(smem_base);
    int* s_cnt = reinterpret_cast<int*>(smem_base + 256);
    float* s_dc = reinterpret_cast<float*>(smem_base + 512);                  // [T][JW_CHAIN_SB]
    // commit list(s): [cap ints][T*cap floats] each.  With B.xcount_smem >= 0 (dedicated chain CTA of the
    // lagged schedule) two lists alternate: this block writes list (ib & 1) and reads the previous
    // block's commits from the other one -- the cross-Gram correction then needs no global list.
    const bool two_lists = B.xcount_smem >= 0;
    unsigned char* lbase = smem_base + 512 + T * JW_CHAIN_SB * 4;
    const size_t lbytes = (size_t)list_cap * 4 * (1 + T);
    int* s_lidx = reinterpret_cast<int*>(lbase + (two_lists ? (size_t)(ib & 1) * lbytes : 0));
    float* s_ld = reinterpret_cast<float*>(s_lidx + list_cap);                // [T][list_cap]
    const int* s_pidx = reinterpret_cast<const int*>(lbase + (size_t)((ib & 1) ^ 1) * lbytes);
    const float* s_pd = reinterpret_cast<const float*>(s_pidx + list_cap);

    const int64_t s = B.b > 0 ? B.s : A.starts[ib];
    const int b = B.b > 0 ? B.b : (int)(A.starts[ib + 1] - s);
    const int SB = (int)blockDim.x;
    const int nsub = (b + SB - 1) / SB;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nw = (int)(blockDim.x >> 5);
    const int64_t p = A.p;
    const float* G = A.gram + (B.b > 0 ? B.gram_off : A.gram_off[ib]);
    unsigned long long my_active = 0, my_rounds = 0;
    unsigned long long ct[6] = {0, 0, 0, 0, 0, 0};
    const bool ctimed = (threadIdx.x == 0) && (A.counters != nullptr) && A.timers;
    unsigned long long ctm = 0;
    if (ctimed) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ctm));
#define JW_CT(i) do { if (ctimed) { unsigned long long n__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n__)); ct[i] += n__ - ctm; ctm = n__; } } while (0)
    int parity = 0;
    int ncommit = 0;          // commits recorded for later sub-blocks (uniform across the CTA)
    int act_total = 0;        // markers with a net delta so far (ordered active list)

    // pull the NEXT block's chain inputs (state, statistics, precomputed terms) towards L2 in bulk, one
    // lane per array: they are needed one whole chain later, so the latency is off the critical path
    if (warp == 0 && B.prefetch_b > 0) {
        const int64_t ps = B.prefetch_s; const int pb = B.prefetch_b;
        for (int kk = 0; kk < T; ++kk) {
            const void* base = nullptr; unsigned esz = 0;
            switch (lane) {
                case 0: base = A.alpha + kk * p + ps; esz = 4; break;
                case 1: base = A.delta + kk * p + ps; esz = 4; break;
                case 2: if (kk == 0) { base = A.xpx + ps; esz = 4; } break;
                case 3: if (kk == 0) { base = A.means + ps; esz = 4; } break;
                case 4: if (METHOD != 1) { base = A.beta + kk * p + ps; esz = 4; } break;
                case 5: if (A.prep_beta0 && kk == 0) { base = A.prep_beta0 + ps; esz = 4; } break;
                case 6: case 7: case 8: case 9: case 10: case 11:
                    if (A.prep && kk == 0) { base = A.prep + (int64_t)(lane - 6) * p + ps; esz = 8; } break;
                default: break;
            }
            if (base) {
                const unsigned long long a0 = (unsigned long long)base & ~15ull;
                const unsigned bytes = (unsigned)((((unsigned long long)base + (unsigned long long)pb * esz + 15ull) & ~15ull) - a0);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
            }
        }
    }

  for (int sb = 0; sb < nsub; ++sb) {
    const int m = sb * SB + tid;           // marker position inside the panel
    const bool valid = m < b;
    const int64_t j = s + (valid ? m : 0);

    // state at block entry
    double r[T];
    float a_entry[T], a_cur[T], b_cur[T];
    int d_cur[T];
    const double mu = (double)A.means[j];
#pragma unroll
    for (int k = 0; k < T; ++k) {
        a_entry[k] = a_cur[k] = A.alpha[k * p + j];
        b_cur[k] = (METHOD == 1) ? 0.0f : A.beta[k * p + j];
        d_cur[k] = A.delta[k * p + j];
    }
    jw_marker_eval<METHOD, T> E;
    E.load_constants(A, j);
    const double x = E.x; (void)x;
    // markers that already carry an effect are certain to need their Gram row: start pulling it
    // towards L2 now (one bulk prefetch per row)
    bool row_requested = false;
    if (valid) {
        bool nz = false;
#pragma unroll
        for (int k = 0; k < T; ++k) nz = nz || (a_cur[k] != 0.0f);
        if (nz) {
            row_requested = true;
            const float* row = G + (int64_t)m * b;
            const unsigned long long a0 = (unsigned long long)row & ~15ull;
            const unsigned bytes = (unsigned)((((unsigned long long)(row + b) + 15ull) & ~15ull) - a0);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
        }
    }

    JW_CT(0);
    if (sb == 0) { if (!wait_fn()) return -1; }
    JW_CT(1);

    // rhs of this marker for every trait
    bool ll_ok = true;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        // .cg loads: these words were produced by other CTAs' atomics in the fused engine
        long long dq, mq, sqk;
        if (MULTI) {
            // exact int64 sums over the ranks' partial rhs (any order gives the same bits)
            dq = 0; mq = 0; sqk = 0;
            ll_ok = jw_ll_rhs(B, T, k, valid ? m : 0, A.mq != nullptr, dq, mq, sqk) && ll_ok;
        } else {
            dq = __ldcg(&A.dq[k * p + j]); mq = A.mq ? __ldcg(&A.mq[k * p + j]) : 0ll;
            sqk = __ldcg(&B.sq[k]);
        }
        r[k] = ((double)dq - mu * (double)(sqk - mq)) * A.invscale;
    }
    if (MULTI) { if (__syncthreads_or(ll_ok ? 0 : 1)) return -1; }
    if (B.xgram != nullptr && valid && two_lists) {
        // previous block's commits from shared memory; four cross-Gram loads in flight, adds in order
        const int xc = B.xcount_smem;
        for (int e0 = 0; e0 < xc; e0 += 4) {
            float g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) g[q] = (e0 + q < xc) ? B.xgram[(int64_t)s_pidx[e0 + q] * b + m] : 0.0f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (e0 + q < xc) {
#pragma unroll
                    for (int k = 0; k < T; ++k) {
                        const float d = s_pd[k * list_cap + e0 + q];
                        if (d != 0.0f) r[k] += (double)d * (double)g[q];
                    }
                }
            }
        }
    } else if (B.xgram != nullptr && valid) {
        const int xc = __ldcg(B.xcount);
        // four entries' loads are in flight together; the additions stay in commit order
        for (int e0 = 0; e0 < xc; e0 += 4) {
            int64_t ja[4]; float g[4]; float dd[4][T];
#pragma unroll
            for (int q = 0; q < 4; ++q) ja[q] = (e0 + q < xc) ? (int64_t)__ldcg(B.xlist + e0 + q) : -1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (ja[q] >= 0) {
                    g[q] = B.xgram[(ja[q] - B.xstart) * b + m];
#pragma unroll
                    for (int k = 0; k < T; ++k) dd[q][k] = __ldcg(&A.dalpha[k * p + ja[q]]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (ja[q] >= 0) {
#pragma unroll
                    for (int k = 0; k < T; ++k)
                        if (dd[q][k] != 0.0f) r[k] += (double)dd[q][k] * (double)g[q];
                }
            }
        }
    }
    if (sb > 0 && valid) {
        for (int e = 0; e < ncommit; ++e) {
            const float g = G[(int64_t)s_lidx[e] * b + m];
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const float d = s_ld[k * list_cap + e];
                if (d != 0.0f) r[k] += (double)d * (double)g;
            }
        }
    }

    JW_CT(2);
    const int nreps = A.nreps_mode ? b : 1;

    for (int rep = 0; rep < nreps; ++rep) {
        E.load_draws(A, j, rep);
        int pos = 0;                   // position inside the sub-block
        while (true) {
            // ---- evaluate this marker against the current rhs ----
            bool pending = valid && tid >= pos;
            float newA[T], newB[T]; int newD[T];
            bool active = false;
            if (pending) {
                active = E.eval(A, r, a_cur, b_cur, d_cur, newA, newB, newD);
                if (active) {
#pragma unroll
                    for (int k = 0; k < T; ++k) s_dc[k * JW_CHAIN_SB + tid] = a_cur[k] - newA[k];
                    if (!row_requested) {
                        // first time this marker looks active: start pulling its Gram row towards L2
                        row_requested = true;
                        const float* row = G + (int64_t)m * b;
                        const unsigned long long a0 = (unsigned long long)row & ~15ull;
                        const unsigned bytes = (unsigned)((((unsigned long long)(row + b) + 15ull) & ~15ull) - a0);
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
                    }
                }
            }
            // ---- first active marker among the pending ones: ONE barrier per round ----
            int key = (pending && active) ? tid : 0x7fffffff;
            int wmin = __reduce_min_sync(0xffffffffu, key);
            if (lane == 0) s_wmin[parity][warp] = wmin;
            __syncthreads();
            int v = (lane < nw) ? s_wmin[parity][lane] : 0x7fffffff;
            const int first = __reduce_min_sync(0xffffffffu, v);
            parity ^= 1;
            my_rounds += (tid == 0);
            // ---- commit everything up to and including `first` ----
            if (pending && tid <= first) {
#pragma unroll
                for (int k = 0; k < T; ++k) { a_cur[k] = newA[k]; b_cur[k] = newB[k]; d_cur[k] = newD[k]; }
                if (tid == first) my_active += 1;
            }
            if (first == 0x7fffffff) break;
            // ---- apply the committed marker's Gram row to the rhs ----
            const int fg = sb * SB + first;        // committed marker's position inside the panel
            if (valid && (A.nreps_mode || tid > first)) {
                float g = G[(int64_t)fg * b + m];
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    float d = s_dc[k * JW_CHAIN_SB + first];
                    if (d != 0.0f) r[k] += (double)d * (double)g;
                }
            }
            if (B.xgram_next != nullptr && tid == 0) {
                // the next block's chain will need this marker's cross-Gram row: start moving it to L2
                const float* row = B.xgram_next + (int64_t)fg * B.b_next;
                const unsigned long long a0 = (unsigned long long)row & ~15ull;
                const unsigned bytes = (unsigned)((((unsigned long long)(row + B.b_next) + 15ull) & ~15ull) - a0);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
            }
            if (sb + 1 < nsub || two_lists) {      // later sub-blocks (and the next block) replay this commit
                if (tid == 0) {
                    s_lidx[ncommit] = fg;
#pragma unroll
                    for (int k = 0; k < T; ++k) s_ld[k * list_cap + ncommit] = s_dc[k * JW_CHAIN_SB + first];
                }
                ncommit += 1;
            }
            pos = first + 1;
            if (pos >= SB || sb * SB + pos >= b) break;
        }
        __syncthreads();      // s_dc / s_wmin are reused by the next repetition
    }

    JW_CT(3);
    // ---- block exit: publish state and the net delta-alpha of every marker ----
    bool any = false;
    if (valid) {
#pragma unroll
        for (int k = 0; k < T; ++k) {
            A.alpha[k * p + j] = a_cur[k];
            if (METHOD != 1) A.beta[k * p + j] = b_cur[k];
            A.delta[k * p + j] = d_cur[k];
            float d = a_entry[k] - a_cur[k];
            A.dalpha[k * p + j] = d;
            any = any || (d != 0.0f);
        }
    }
    if (B.write_active_list) {
        // ordered compaction of this sub-block's markers with any non-zero delta (nothing to do in the
        // common case of a sub-block without updates)
        if (__syncthreads_or(any ? 1 : 0)) {
            unsigned bal = __ballot_sync(0xffffffffu, any);
            if (lane == 0) s_cnt[warp] = __popc(bal);
            __syncthreads();
            int c = (lane < nw) ? s_cnt[lane] : 0;          // every warp scans the 32 warp counts
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int warp_off = __shfl_sync(0xffffffffu, incl - c, warp);
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (any) B.act_idx[act_total + warp_off + __popc(bal & ((1u << lane) - 1u))] = (int32_t)j;
            act_total += total;
        }
    }
    __syncthreads();          // shared scratch is reused by the next sub-block
    JW_CT(4);
  }   // sub-blocks
    if (B.write_active_list && tid == 0) *B.act_cnt = act_total;
    if (ctimed) for (int i = 0; i < 5; ++i) atomicAdd(&A.counters[56 + i], ct[i]);
#undef JW_CT
    if (A.counters) {
        if (my_active) atomicAdd(&A.counters[0], my_active);
        if (my_rounds) atomicAdd(&A.counters[1], my_rounds);
    }
    return ncommit;
}

template <int METHOD, int T>
__global__ void __launch_bounds__(JW_MAX_BLOCK)
jw_k_chain(jw_chain_args A, int list_cap) {
    extern __shared__ __align__(16) unsigned char jw_chain_dyn[];
    jw_chain_block<METHOD, T, false>(A, jw_chain_blk_from(A), A.block0 + (int)blockIdx.x, jw_no_wait(), jw_chain_dyn, list_cap);
}

// ------------------------------------------------------------------------------------------
// ordered compaction of all markers with a non-zero delta (independent-block reconcile,
// BayesABC.jl:251-253): single CTA, contiguous chunk per thread, so the list is ascending.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
jw_k_compact_active(const float* __restrict__ dalpha, int64_t p, int t,
                    int32_t* __restrict__ act_idx, int32_t* __restrict__ act_cnt) {
    __shared__ int s_cnt[1025];
    int64_t chunk = (p + 1023) / 1024;
    int64_t beg = (int64_t)threadIdx.x * chunk, end = min(beg + chunk, p);
    int c = 0;
    for (int64_t j = beg; j < end; ++j) {
        bool any = false;
        for (int k = 0; k < t; ++k) any = any || (dalpha[k * p + j] != 0.0f);
        c += any;
    }
    s_cnt[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < 1024; ++i) { int v = s_cnt[i]; s_cnt[i] = acc; acc += v; }
        s_cnt[1024] = acc;
        *act_cnt = acc;
    }
    __syncthreads();
    int o = s_cnt[threadIdx.x];
    for (int64_t j = beg; j < end; ++j) {
        bool any = false;
        for (int k = 0; k < t; ++k) any = any || (dalpha[k * p + j] != 0.0f);
        if (any) act_idx[o++] = (int32_t)j;
    }
}

// ------------------------------------------------------------------------------------------
// block exit  ycorr[k] += X_b * dalpha[k]  (BayesABC.jl:181-185): one thread per individual,
// active columns applied in ascending marker order with one fused multiply-add each, exactly as
// the oracle does.  x = code - mean, 0 where the call is missing (decode_marker!).
// ------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(256)
jw_k_apply(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t n, int64_t p,
           const float* __restrict__ means, const float* __restrict__ dalpha,
           const int32_t* __restrict__ act_idx, const int32_t* __restrict__ act_cnt,
           float* __restrict__ y, int64_t r0, int64_t r1) {
    int64_t i = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = *act_cnt;
    if (i >= r1 || cnt == 0) return;
    float v[T];
#pragma unroll
    for (int k = 0; k < T; ++k) v[k] = y[k * n + i];
    const int sh = (int)(i & 3) << 1;
    const int64_t byte = i >> 2;
    for (int a = 0; a < cnt; ++a) {
        int64_t j = act_idx[a];
        unsigned code = (packed[j * stride_d + byte] >> sh) & 3u;
        float mu = means[j];
        float xv = (code == 3u ? mu : (float)code) - mu;
#pragma unroll
        for (int k = 0; k < T; ++k) {
            float d = dalpha[k * p + j];
            if (d != 0.0f) v[k] = fmaf(d, xv, v[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < T; ++k) y[k * n + i] = v[k];
}

// out = M * alpha (getEBV / ycorr init): same column walk, alpha as the coefficient, no list
// (rows [r0, r1) of this rank; `packed` is addressed by global row)
__global__ void __launch_bounds__(256)
jw_k_mul_alpha(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t r0, int64_t r1, int64_t p,
               const float* __restrict__ means, const float* __restrict__ alpha,
               float sign, float* __restrict__ out, int accumulate) {
    int64_t i = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r1) return;
    float v = accumulate ? out[i] : 0.0f;
    const int sh = (int)(i & 3) << 1;
    const int64_t byte = i >> 2;
    for (int64_t j = 0; j < p; ++j) {
        float a = alpha[j];
        if (a != 0.0f) {
            unsigned code = (packed[j * stride_d + byte] >> sh) & 3u;
            float mu = means[j];
            float xv = (code == 3u ? mu : (float)code) - mu;
            v = fmaf(sign * a, xv, v);
        }
    }
    out[i] = v;
}

// ------------------------------------------------------------------------------------------
// canonical reductions: chunks of 256 consecutive elements summed in index order by one thread,
// chunk sums then added in order by a single thread.  Same order in tests/ (canonical_sum).
// ------------------------------------------------------------------------------------------
#define JW_CHUNK 256
__global__ void __launch_bounds__(128)
jw_k_chunk_prod(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                double* __restrict__ partials) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t beg = c * JW_CHUNK;
    if (beg >= n) return;
    int64_t end = min(beg + JW_CHUNK, n);
    double s = 0.0;
    for (int64_t i = beg; i < end; ++i) s += (double)a[i] * (double)b[i];
    partials[c] = s;
}
__global__ void __launch_bounds__(128)
jw_k_chunk_sum(const float* __restrict__ a, int64_t n, double* __restrict__ partials) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t beg = c * JW_CHUNK;
    if (beg >= n) return;
    int64_t end = min(beg + JW_CHUNK, n);
    double s = 0.0;
    for (int64_t i = beg; i < end; ++i) s += (double)a[i];
    partials[c] = s;
}
// chunk sums -> sums of 256 consecutive chunk sums (index order, one thread each) -> final sum of those
// (index order, one thread).  Same three levels in tests/helpers.py:canonical_sum_prod.
__global__ void __launch_bounds__(JW_CHUNK)
jw_k_chunk_final(const double* __restrict__ partials, int64_t nchunks, double* out) {
    __shared__ double s_g[JW_CHUNK];
    const int64_t ngroups = (nchunks + JW_CHUNK - 1) / JW_CHUNK;      // <= 256 for up to 16.7M elements
    double s = 0.0;
    if (threadIdx.x < ngroups) {
        const int64_t beg = (int64_t)threadIdx.x * JW_CHUNK, end = min(beg + JW_CHUNK, nchunks);
        for (int64_t c = beg; c < end; ++c) s += partials[c];
    }
    s_g[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int64_t g = 0; g < ngroups; ++g) tot += s_g[g];
        *out = tot;
    }
}
// BayesR: sum alpha^2 / gamma[delta] over delta > 1 (variance_components.jl:68-79)
__global__ void __launch_bounds__(128)
jw_k_chunk_bayesr(const float* __restrict__ alpha, const int32_t* __restrict__ delta, int64_t n,
                  double g1, double g2, double g3, double g4, double g5, double g6, double g7,
                  double* __restrict__ partials) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t beg = c * JW_CHUNK;
    if (beg >= n) return;
    int64_t end = min(beg + JW_CHUNK, n);
    double g[8] = {0.0, g1, g2, g3, g4, g5, g6, g7};
    double s = 0.0;
    for (int64_t i = beg; i < end; ++i) {
        int d = delta[i];
        if (d > 1) s += (double)alpha[i] * (double)alpha[i] / g[d - 1];
    }
    partials[c] = s;
}
// integer statistics: counts are exact under any order.  Per-thread flags are reduced inside the block
// (ballots + shared counters) so that each block issues one atomic per counter, not one per marker.
__global__ void __launch_bounds__(256)
jw_k_counts(const float* __restrict__ alpha, const int32_t* __restrict__ delta, int64_t p, int t,
            int method, unsigned long long* __restrict__ out /* [0..3] nnz, [4..7] sumdelta, [8..23] classes */) {
    __shared__ unsigned int s_c[24];
    if (threadIdx.x < 24) s_c[threadIdx.x] = 0;
    __syncthreads();
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = j < p;
    int state = 0, cls = -1;
    for (int k = 0; k < t; ++k) {
        const bool nz = in && alpha[k * p + j] != 0.0f;
        const int d = in ? delta[k * p + j] : 0;
        const bool dl = in && (method == 1 ? d > 1 : d != 0);
        unsigned b1 = __ballot_sync(0xffffffffu, nz), b2 = __ballot_sync(0xffffffffu, dl);
        if ((threadIdx.x & 31) == 0) {
            if (b1) atomicAdd(&s_c[k], __popc(b1));
            if (b2) atomicAdd(&s_c[4 + k], __popc(b2));
        }
        state |= (d != 0) << k;
        if (k == 0) cls = d - 1;
    }
    const int bin = method == 1 ? cls : state;
    for (int c = 0; c < 16; ++c) {
        unsigned bb = __ballot_sync(0xffffffffu, in && bin == c);
        if ((threadIdx.x & 31) == 0 && bb) atomicAdd(&s_c[8 + c], __popc(bb));
    }
    __syncthreads();
    if (threadIdx.x < 24 && s_c[threadIdx.x]) atomicAdd(&out[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}
// sum / max|.| of ycorr (order-free: integer-like max; the sum is informational)
__global__ void __launch_bounds__(256)
jw_k_maxabs(const float* __restrict__ y, int64_t n, unsigned* __restrict__ out) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(y[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
// all-gather staging: recv = [rank][trait][chunk] -> y[trait][bounds[rank] + i]
struct jw_bounds { int64_t b[9]; };
__global__ void __launch_bounds__(256)
jw_k_scatter_rows(const float* __restrict__ recv, int64_t chunk, int t, int world, jw_bounds B,
                  float* __restrict__ y, int64_t n, int skip_rank) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y / t, k = blockIdx.y % t;
    if (r == skip_rank) return;
    const int64_t len = B.b[r + 1] - B.b[r];
    if (i < len) y[(int64_t)k * n + B.b[r] + i] = recv[((int64_t)r * t + k) * chunk + i];
}
__global__ void __launch_bounds__(256)
jw_k_shift(float* __restrict__ y, int64_t n, float shift) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = y[i] + shift;
}
__global__ void __launch_bounds__(256)
jw_k_fill_double(double* __restrict__ a, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// posterior running means (output.jl:568-577): Float32 state, binary64 update
__global__ void __launch_bounds__(256)
jw_k_accumulate(const float* __restrict__ alpha, const int32_t* __restrict__ delta, int64_t tp,
                double nsamples, int bayesr, float* __restrict__ ma, float* __restrict__ ma2,
                float* __restrict__ md) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tp) return;
    double a = (double)alpha[i];
    double d = bayesr ? (delta[i] > 1 ? 1.0 : 0.0) : (double)delta[i];
    ma[i] = (float)((double)ma[i] + (a - (double)ma[i]) / nsamples);
    ma2[i] = (float)((double)ma2[i] + (a * a - (double)ma2[i]) / nsamples);
    md[i] = (float)((double)md[i] + (d - (double)md[i]) / nsamples);
}

// BayesB: var_j = (beta_j^2 + df*scale)/chisq(df+1) (variance_components.jl:60-66, 169-172).
// chisq(k) = 2*Gamma(k/2) by Marsaglia-Tsang with draws from the native stream
// (pseudo-traits 126/127 of the marker's counter space, attempt number in `rep`).
__global__ void __launch_bounds__(256)
jw_k_bayesb_var(const float* __restrict__ beta, int64_t p, double df, double scale,
                uint64_t seed, uint32_t iter, double* __restrict__ ve) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    double shape = 0.5 * (df + 1.0);        // >= 0.5; boost when < 1
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
