// jw_setup_kernels.cuh -- one-time kernels: per-marker statistics and Gram blocks from the
// 2-bit packed image.  Replaces GibbsMats (markers/tools4genotypes.jl:237-275): xpRinvx by
// getXpRinvX (:28-36) and XpRinvX = Xblock'Xblock (:263).  Everything is integer popcount
// work followed by one binary64 closed form per entry (see gram_value in the oracle).
#pragma once
#include "jw_common.cuh"

// centred cross-product from integer sufficient statistics; must match
// oracle/jwas_oracle.c:gram_value operation for operation.
__device__ __forceinline__ float jw_gram_value(long long Nab, long long Sa_vb, long long Sb_va,
                                               long long Nvv, float mua, float mub) {
    double ma = (double)mua, mb = (double)mub;
    double g = (double)Nab - ma * (double)Sb_va;
    g = g - mb * (double)Sa_vb;
    g = g + (ma * mb) * (double)Nvv;
    return (float)g;
}

// one warp per marker: counts of codes 1, 2 and 3 over the rows this rank stores (zero padding counts nothing)
__global__ void __launch_bounds__(256)
jw_k_marker_counts(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t p, int32_t* __restrict__ cnt) {
    int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (j >= p) return;
    const uint32_t* col = reinterpret_cast<const uint32_t*>(packed + j * stride_d);
    int64_t nwords = stride_d >> 2;
    int n1 = 0, n2 = 0, nm = 0;
    for (int64_t w = lane; w < nwords; w += 32) {
        uint32_t v = __ldg(col + w);
        uint32_t lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
        n1 += __popc(lo & ~hi);
        n2 += __popc(hi & ~lo);
        nm += __popc(lo & hi);
    }
    for (int o = 16; o > 0; o >>= 1) {
        n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        nm += __shfl_xor_sync(0xffffffffu, nm, o);
    }
    if (lane == 0) { cnt[j] = n1; cnt[p + j] = n2; cnt[2 * p + j] = nm; }
}

// counts (summed over the ranks when rows are sharded) -> mean, xpx, colsum, nvalid.  ext_means: the means
// were supplied by the host (centring on a larger sample than the analysed rows, readgenotypes.jl:372-385
// before the alignment of JWAS.jl:381-402); xpx follows from the same closed form.
__global__ void __launch_bounds__(256)
jw_k_marker_finalize(const int32_t* __restrict__ cnt, int64_t n, int64_t p, int ext_means,
                     float* __restrict__ means, float* __restrict__ xpx,
                     int32_t* __restrict__ colsum, int32_t* __restrict__ nvalid, int* __restrict__ has_missing) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const int n1 = cnt[j], n2 = cnt[p + j], nm = cnt[2 * p + j];
    long long nn = (long long)n - nm;
    long long sum = (long long)n1 + 2ll * n2;
    float mu = ext_means ? means[j] : (nn > 0 ? (float)sum / (float)nn : 0.0f);
    means[j] = mu;
    xpx[j] = jw_gram_value((long long)n1 + 4ll * n2, sum, sum, nn, mu, mu);
    colsum[j] = (int32_t)sum;
    nvalid[j] = (int32_t)nn;
    if (nm > 0) atomicOr(has_missing, 1);
}

// Gram blocks.  One CTA computes a 64x64 tile of one block's b*b matrix, streaming both
// operand panels through shared memory 32 words (=512 individuals) at a time; each of the
// 256 threads owns a 4x4 micro-tile of integer pair counts.
#define JW_GT 64
#define JW_GK 32
template <bool MISSING>
__global__ void __launch_bounds__(256)
jw_k_gram(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t n,
          const float* __restrict__ means, const int32_t* __restrict__ colsum,
          const int64_t* __restrict__ starts, const int64_t* __restrict__ tile_off,
          const int32_t* __restrict__ tile_blocks, const int32_t* __restrict__ tile_ab,
          float* __restrict__ gram) {
    __shared__ uint32_t sA[JW_GT][JW_GK + 1];
    __shared__ uint32_t sB[JW_GT][JW_GK + 1];
    // rows come from marker block rb, columns from block cb (rb == cb: the block's own Gram;
    // cb == rb + 1: the cross-Gram used by the lagged schedule)
    const int rb = tile_blocks[2 * blockIdx.x], cb = tile_blocks[2 * blockIdx.x + 1];
    const int ta = tile_ab[2 * blockIdx.x], tb = tile_ab[2 * blockIdx.x + 1];
    const int64_t s = starts[rb], sc = starts[cb];
    const int b = (int)(starts[rb + 1] - s), bc = (int)(starts[cb + 1] - sc);
    const int a0 = ta * JW_GT, c0 = tb * JW_GT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // micro-tile coordinates
    const int64_t nwords = stride_d >> 2;
    int Nab[4][4];
    int Sa[4][4], Sb[4][4], Nvv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) { Nab[i][k] = 0; Sa[i][k] = 0; Sb[i][k] = 0; Nvv[i][k] = 0; }

    for (int64_t w0 = 0; w0 < nwords; w0 += JW_GK) {
        // cooperative, coalesced load of 64 x 32 words for each operand
        for (int e = threadIdx.x; e < JW_GT * JW_GK; e += 256) {
            int r = e / JW_GK, c = e % JW_GK;
            int64_t w = w0 + c;
            uint32_t va = 0xffffffffu, vb = 0xffffffffu;   // out of range = missing: counts nothing
            if (w < nwords) {
                if (a0 + r < b) va = __ldg(reinterpret_cast<const uint32_t*>(packed + (s + a0 + r) * stride_d) + w);
                if (c0 + r < bc) vb = __ldg(reinterpret_cast<const uint32_t*>(packed + (sc + c0 + r) * stride_d) + w);
            }
            sA[r][c] = va; sB[r][c] = vb;
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < JW_GK; ++c) {
            uint32_t a1[4], a2[4], av[4], b1[4], b2[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t v = sA[ty * 4 + i][c];
                uint32_t lo = v & 0x55555555u, hi = (v >> 1) & 0x55555555u;
                a1[i] = lo & ~hi; a2[i] = hi & ~lo; av[i] = ~(lo & hi) & 0x55555555u;
                v = sB[tx * 4 + i][c];
                lo = v & 0x55555555u; hi = (v >> 1) & 0x55555555u;
                b1[i] = lo & ~hi; b2[i] = hi & ~lo; bv[i] = ~(lo & hi) & 0x55555555u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // codes occupy even bit positions only, so disjoint sets can share one popc
                    Nab[i][k] += __popc(a1[i] & b1[k]) + 2 * __popc((a1[i] & b2[k]) | (a2[i] & b1[k]))
                               + 4 * __popc(a2[i] & b2[k]);
                    if (MISSING) {
                        Sa[i][k] += __popc(a1[i] & bv[k]) + 2 * __popc(a2[i] & bv[k]);
                        Sb[i][k] += __popc(b1[k] & av[i]) + 2 * __popc(b2[k] & av[i]);
                        Nvv[i][k] += __popc(av[i] & bv[k]);
                    }
                }
        }
        __syncthreads();
    }
    const long long pad = (long long)nwords * 16 - n;   // zero padding reads as "observed code 0"
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int a = a0 + ty * 4 + i, c = c0 + tx * 4 + k;
            if (a < b && c < bc) {
                long long sa, sb, nvv;
                if (MISSING) { sa = Sa[i][k]; sb = Sb[i][k]; nvv = (long long)Nvv[i][k] - pad; }
                else { sa = colsum[s + a]; sb = colsum[sc + c]; nvv = n; }
                gram[tile_off[blockIdx.x] + (int64_t)a * bc + c] =
                    jw_gram_value(Nab[i][k], sa, sb, nvv, means[s + a], means[sc + c]);
            }
        }
}


// ------------------------------------------------------------------------------------------
// Gram blocks on the tensor cores.  X_b'X_b IS a GEMM in the reference (tools4genotypes.jl:263,
// `Xblock' * Xblock` -> sgemm), so the plain library GEMM is the right tool: codes 0/1/2 and the
// observed-mask 0/1 are exact in bf16, products are exact, and FP32 accumulation of integers below
// 2^24 is exact in any order -- the integer pair counts come out bit-identical to the popcount
// kernel above, which stays as the fall-back (n >= 2^22) and as the cross-check in the tests.
// ------------------------------------------------------------------------------------------
#include <cuda_bf16.h>

// unpack the markers [s, s+b) into column-major bf16 matrices (lda = n_pad): C = code value (missing -> 0),
// V = 1 where observed (only written when MISSING)
template <bool MISSING>
__global__ void __launch_bounds__(256)
jw_k_unpack_bf16(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t n, int64_t n_pad,
                 int64_t s, int b, __nv_bfloat16* __restrict__ C, __nv_bfloat16* __restrict__ V) {
    const int64_t nbytes_pad = n_pad >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (marker, byte)
    if (idx >= (int64_t)b * nbytes_pad) return;
    const int m = (int)(idx / nbytes_pad); const int64_t byte = idx % nbytes_pad;
    const unsigned v = (byte < stride_d) ? packed[(s + m) * stride_d + byte] : 0u;
    __nv_bfloat16 c[4], o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = byte * 4 + k;
        unsigned code = (v >> (2 * k)) & 3u;
        const bool obs = (code != 3u) && (i < n);
        c[k] = __float2bfloat16((obs ? (float)code : 0.0f));
        o[k] = __float2bfloat16(obs ? 1.0f : 0.0f);
    }
    __nv_bfloat16* dst = C + (int64_t)m * n_pad + byte * 4;
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(c);
    if (MISSING) *reinterpret_cast<uint2*>(V + (int64_t)m * n_pad + byte * 4) = *reinterpret_cast<uint2*>(o);
}

// exact integer counts (as floats) -> centred Float32 Gram values
template <bool MISSING>
__global__ void __launch_bounds__(256)
jw_k_gram_finalize(const float* __restrict__ Nab, const float* __restrict__ Sa, const float* __restrict__ Sb,
                   const float* __restrict__ Nvv, int64_t n, const float* __restrict__ means,
                   const int32_t* __restrict__ colsum, int64_t s_r, int b_r, int64_t s_c, int b_c,
                   float* __restrict__ out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)b_r * b_c) return;
    const int a = (int)(idx / b_c), c = (int)(idx % b_c);
    long long sa, sb, nvv;
    if (MISSING) { sa = (long long)Sa[idx]; sb = (long long)Sb[idx]; nvv = (long long)Nvv[idx]; }
    else { sa = colsum[s_r + a]; sb = colsum[s_c + c]; nvv = n; }
    out[idx] = jw_gram_value((long long)Nab[idx], sa, sb, nvv, means[s_r + a], means[s_c + c]);
}
