// jw_stream_kernel.cuh -- the block rhs of ALL marker blocks against one ycorr snapshot, as a single streamed pass over
// the tiled genotype image: the GEMV of the independent-block schedule (BayesABC_block_independent! computes every
// block's rhs from yCorr_snapshot, BayesABC.jl:205-217; BayesR.jl:207-222; MTBayesABC.jl:350-366).
//
// Same arithmetic and the same lookup-table stream as the persistent sweep kernel (jw_fused_sweep.cuh: one table
// lookup + one integer add per packed byte = 4 genotypes), but with no chain on the critical path: a CTA quantises its
// row slice and builds its tables ONCE, then its warps run through every 16-marker chunk of every block without a
// barrier.  Per-panel fixed costs (record replay, re-quantisation, table rebuild, end-of-panel barrier) do not exist
// here, so this kernel shows what the lookup-table stream itself sustains.
//   NT  threads per CTA (1 CTA per SM: the tables take 128-192 KB of shared memory)
//   DB  1 = the next chunk's 128-bit loads are issued before the current chunk is reduced (register double buffer)
#pragma once
#include "jw_common.cuh"
#include "jw_fused_sweep.cuh"

struct jw_stream_args {
    const uint8_t* tiled; const int64_t* chunk_off; const int32_t* chunk_block; const int64_t* starts;
    int Gs, n_vs, nblocks, uniform_b;
    int64_t total_chunks;
    int64_t n, p, row_off, nloc;
    const float* ycorr; float scale;
    long long* dq; long long* mq; long long* sq;
    int32_t* flags;
    int pf_dist;                  // > 0: each warp pulls the chunk it will process pf_dist iterations later into L2
};

template <int T, int W, int NT, int DB>
__global__ void __launch_bounds__(NT, 1)
jw_k_stream(jw_stream_args S) {
    extern __shared__ __align__(16) int jw_smem[];
    constexpr bool MISS = (T == 1 && W == 2);
    constexpr int NGB = JW_FUSED_MAX_GS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = NT / 32;
    const int Gs = S.Gs, R = Gs * 4;
    unsigned char* tab = reinterpret_cast<unsigned char*>(jw_smem);
    int* yqs = jw_smem + JW_TAB_BYTES(W) / 4;            // [T][R]
    __shared__ long long s_red[32 * JW_MAX_TRAITS];
    const int64_t n = S.n, p = S.p;
    const float* const ycorr_l = S.ycorr + S.row_off;
    const int cpb = (S.uniform_b + 15) >> 4;             // chunks per block when every block has uniform_b markers

    for (int vs = blockIdx.x; vs < S.n_vs; vs += gridDim.x) {
        const int64_t row0 = (int64_t)vs * R;
        // ---- fixed-point image of the slice (once per sweep) ----
        long long qs[T];
#pragma unroll
        for (int kk = 0; kk < T; ++kk) qs[kk] = 0;
        for (int r = tid; r < R; r += NT) {
            const int64_t row = row0 + r;
            const bool rv = row < S.nloc;
            int ovf = 0;
#pragma unroll
            for (int kk = 0; kk < T; ++kk) {
                const int q = rv ? jw_quantize(ycorr_l[kk * n + row], S.scale, &ovf) : 0;
                yqs[kk * R + r] = q;
                qs[kk] += q;
            }
            if (ovf) atomicOr(&S.flags[0], 1);
        }
#pragma unroll
        for (int kk = 0; kk < T; ++kk) {
            long long v = qs[kk];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[warp * JW_MAX_TRAITS + kk] = v;
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < T; ++kk) {
                long long v = 0;
                for (int w = 0; w < nwarps; ++w) v += s_red[w * JW_MAX_TRAITS + kk];
                if (v != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&S.sq[kk]), (unsigned long long)v);
            }
        }
        // ---- lookup tables: entry e of group g = sum over its 4 individuals (layout: jw_fused_sweep.cuh) ----
        for (int item = tid; item < Gs * 16; item += NT) {
            const int g = item % Gs, ehi = item / Gs;
            const unsigned c2 = ehi & 3, c3 = ehi >> 2;
#pragma unroll
            for (int comp = 0; comp < W; ++comp) {
                const int tr = MISS ? 0 : comp;
                const bool mc_ = MISS && comp == 1;
                const int y0 = yqs[tr * R + 4 * g], y1 = yqs[tr * R + 4 * g + 1],
                          y2 = yqs[tr * R + 4 * g + 2], y3 = yqs[tr * R + 4 * g + 3];
                const int B = jw_tabval(c2, y2, mc_) + jw_tabval(c3, y3, mc_);
#pragma unroll
                for (int elo = 0; elo < 16; ++elo) {
                    const int val = jw_tabval(elo & 3, y0, mc_) + jw_tabval(elo >> 2, y1, mc_) + B;
                    const int gb = g >> 5, l = g & 31;
                    const int sub = gb == 0 ? jw_tab_sub<W>(0) : (gb == 1 ? jw_tab_sub<W>(1) : jw_tab_sub<W>(2));
                    *reinterpret_cast<int*>(tab + (ehi * 16 + elo) * 256 + sub + l * 4 * W + comp * 4) = val;
                }
            }
        }
        __syncthreads();

        // ---- the stream: chunk gc = 16 markers of one block; warps take consecutive chunks.  (k, mc) = (block, chunk
        //      inside the block) advance incrementally: no division in the loop ----
        struct cursor { int gc, k, mc; };
        auto advance = [&](cursor& c, const int step) {
            c.gc += step;
            if (S.uniform_b > 0) { c.mc += step; while (c.mc >= cpb) { c.mc -= cpb; c.k += 1; } }
        };
        auto locate = [&](const cursor& c, int64_t& s, int& b, const uint8_t*& src) {
            int64_t co; int k = c.k, mc = c.mc;
            if (S.uniform_b > 0) {
                co = (int64_t)k * cpb;
                s = (int64_t)k * S.uniform_b; b = (int)min((int64_t)S.uniform_b, p - s);
            } else {
                k = S.chunk_block[c.gc]; co = S.chunk_off[k]; mc = (int)(c.gc - co);
                s = S.starts[k]; b = (int)(S.starts[k + 1] - s);
            }
            const int nch = (b + 15) >> 4;
            s += (int64_t)mc * 16; b -= mc * 16;             // first marker of the chunk, markers left from there
            src = S.tiled + (((size_t)(co * S.n_vs + (int64_t)vs * nch) * Gs + (size_t)mc * Gs) << 4);
        };
        auto load = [&](const uint8_t* src, uint4 (&dv)[NGB]) {
#pragma unroll
            for (int gb = 0; gb < NGB; ++gb) {
                const int g = gb * 32 + lane;
                dv[gb] = make_uint4(0, 0, 0, 0);
                if (g < Gs) dv[gb] = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)g << 4)));
            }
        };
        uint4 dv[NGB], dn[NGB];
        int64_t s_c = 0, s_n = 0; int b_c = 0, b_n = 0;
        const uint8_t* src = nullptr;
        const int total = (int)S.total_chunks;
        cursor cur = {0, 0, 0}, nxt, pfc;
        advance(cur, warp);
        nxt = cur; pfc = cur;
        if (S.pf_dist > 0) advance(pfc, S.pf_dist * nwarps);
        if (DB && cur.gc < total) { locate(cur, s_n, b_n, src); load(src, dn); }
        for (; cur.gc < total; advance(cur, nwarps)) {
            if (DB) {
                s_c = s_n; b_c = b_n;
#pragma unroll
                for (int gb = 0; gb < NGB; ++gb) dv[gb] = dn[gb];
                nxt = cur; advance(nxt, nwarps);
                if (nxt.gc < total) { locate(nxt, s_n, b_n, src); load(src, dn); }
            } else {
                locate(cur, s_c, b_c, src); load(src, dv);
            }
            if (S.pf_dist > 0) {
                if (lane == 0 && pfc.gc < total) {
                    int64_t sp; int bp; const uint8_t* psrc;
                    locate(pfc, sp, bp, psrc);
                    jw_prefetch_l2(psrc, (unsigned)Gs << 4);
                }
                advance(pfc, nwarps);
            }
            int acc[16][W];
#pragma unroll
            for (int q = 0; q < 16; ++q)
#pragma unroll
                for (int comp = 0; comp < W; ++comp) acc[q][comp] = 0;
#pragma unroll
            for (int gb = 0; gb < NGB; ++gb) {
                const int g = gb * 32 + lane;
                if (g < Gs) {
                    const unsigned char* tg = tab + jw_tab_sub<W>(gb);
                    const uint32_t laneoff = (uint32_t)lane * 4u * W;
                    const uint32_t wds[4] = {dv[gb].x, dv[gb].y, dv[gb].z, dv[gb].w};
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const uint32_t off = __byte_perm(wds[q >> 2], laneoff, 0x6504u | ((q & 3) << 4));
                        if (W == 1) {
                            acc[q][0] += *reinterpret_cast<const int*>(tg + off);
                        } else {
                            const int2 e2 = *reinterpret_cast<const int2*>(tg + off);
                            acc[q][0] += e2.x; acc[q][W - 1] += e2.y;
                        }
                    }
                }
            }
            // transposed butterfly: 16 markers x 32 lanes -> marker (lane>>1)&15 on every lane
#pragma unroll
            for (int comp = 0; comp < W; ++comp) {
                long long vals[8];
                {
                    const bool up = (lane & 16) != 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int keep = up ? acc[i + 8][comp] : acc[i][comp];
                        const int send = up ? acc[i][comp] : acc[i + 8][comp];
                        vals[i] = (long long)keep + (long long)__shfl_xor_sync(0xffffffffu, send, 16);
                    }
                }
#pragma unroll
                for (int half = 4, mask = 8; half >= 1; half >>= 1, mask >>= 1) {
                    const bool up = (lane & mask) != 0;
#pragma unroll
                    for (int i = 0; i < half; ++i) {
                        const long long keep = up ? vals[i + half] : vals[i];
                        const long long send = up ? vals[i] : vals[i + half];
                        vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
                    }
                }
                const long long tot = vals[0] + __shfl_xor_sync(0xffffffffu, vals[0], 1);
                const int q = (lane >> 1) & 15;
                if ((lane & 1) == 0 && q < b_c && tot != 0) {
                    long long* dst = (MISS && comp == 1) ? &S.mq[s_c + q] : &S.dq[(int64_t)(MISS ? 0 : comp) * p + s_c + q];
                    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(dst), "l"((unsigned long long)tot) : "memory");
                }
            }
        }
        __syncthreads();           // the next slice overwrites the tables
    }
}

// ------------------------------------------------------------------------------------------ host side
template <int T, int W, int NT, int DB>
static int jw_stream_launch_v(jwas_handle* h, jw_fused_state* f, jw_stream_args& S) {
    auto kern = jw_k_stream<T, W, NT, DB>;
    const size_t smem = (size_t)JW_TAB_BYTES(W) + (size_t)T * f->Gs * 4 * 4 + 64;
    JW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min<int>(h->sm_count, f->n_vs);
    kern<<<grid, NT, smem, h->stream>>>(S);
    h->launches += 1;
    JW_CUDA(cudaGetLastError());
    return 0;
}
template <int T, int W>
static int jw_stream_launch_tw(jwas_handle* h, jw_fused_state* f, jw_stream_args& S) {
    switch ((int)h->opt_stream_variant) {
        case 0: return jw_stream_launch_v<T, W, 512, 1>(h, f, S);
        case 2: return jw_stream_launch_v<T, W, 768, 1>(h, f, S);
        case 3: return jw_stream_launch_v<T, W, 1024, 1>(h, f, S);
        default: return jw_stream_launch_v<T, W, 1024, 0>(h, f, S);
    }
}

static bool jw_stream_supported(const jwas_handle* h) {
    jw_fused_state* f = (jw_fused_state*)h->fused;
    return f && f->ready && (h->t == 1 || (h->t == 2 && !h->has_missing));
}

// dq/mq/sq of every marker against the current ycorr; the caller has zeroed dq, mq and sq
static int jw_stream_all(jwas_handle* h, float scale) {
    jw_fused_state* f = (jw_fused_state*)h->fused;
    JW_REQUIRE(jw_stream_supported(h), "streamed block rhs: unsupported trait/missing combination");
    jw_stream_args S;
    S.tiled = f->d_tiled; S.chunk_off = f->d_chunk_off; S.chunk_block = f->d_chunk_block; S.starts = h->d_starts;
    S.Gs = f->Gs; S.n_vs = f->n_vs; S.nblocks = (int)h->nblocks; S.total_chunks = f->total_chunks;
    S.uniform_b = 0;
    {
        const int64_t b0 = h->starts[1] - h->starts[0];
        bool uni = true;
        for (int64_t k = 1; k < h->nblocks && uni; ++k) {
            const int64_t bk = h->starts[k + 1] - h->starts[k];
            uni = (k + 1 < h->nblocks) ? (bk == b0) : (bk <= b0);
        }
        if (uni) S.uniform_b = (int)b0;
    }
    S.n = h->n; S.p = h->p; S.row_off = h->row_begin; S.nloc = jw_nloc(h);
    S.ycorr = h->d_ycorr; S.scale = scale;
    S.dq = h->d_dq; S.mq = h->d_mq; S.sq = h->d_sq; S.flags = h->d_flags;
    S.pf_dist = (int)h->opt_stream_pf;
    if (h->t == 1 && !h->has_missing) return jw_stream_launch_tw<1, 1>(h, f, S);
    if (h->t == 1) return jw_stream_launch_tw<1, 2>(h, f, S);
    return jw_stream_launch_tw<2, 2>(h, f, S);
}
