// jw_fused_ws.cuh -- streaming role of the persistent sweep kernel, WARP-SPECIALISED (kernel MODE 3).
//
// The plain streaming role (jw_fused_sweep.cuh, MODE 1) runs every panel as  replay the commit records -> barrier ->
// quantise + rebuild the lookup tables -> barrier -> stream -> barrier : ~4 us of every ~13 us panel (cfg2) go to the
// serial part and to the barrier tails.  Here the two activities run CONCURRENTLY on different warps of the CTA with
// two sets of lookup tables in shared memory:
//   * builder warps (JW_WS_NB of the 32): for panel k+1 -- replay the records of block (k+1)-1-lag on the CTA's rows of
//     ycorr, re-quantise, build table set (k+1)&1, publish it through an mbarrier (full[set]);
//   * streaming warps (the other 32 - JW_WS_NB): for panel k -- wait full[k&1], stream the panel (one table lookup +
//     one integer add per packed byte, as everywhere), then EACH WARP by itself releases its part of the panel
//     (red.release.gpu on the panel's arrival counter: no CTA-wide barrier at the end of a panel) and frees the table
//     set (empty[set]) for panel k+2.
// With the lagged schedule (lag >= 1; 2 by default) the records a builder needs were committed a panel or two ago, so it
// practically never waits, and the streaming warps never stop except for a table set that is not ready yet.
// Shared memory (T = 1, no missing calls; 64-bit table entries would need twice the space):
//   A0 | A1 : 64 KB each -- table rows of 256 B holding byte-groups 0..63 (as in MODE 1) of set 0 / set 1
//   B       : 64 KB      -- byte-groups 64..95 of BOTH sets, interleaved inside the 256-byte rows (set * 128 + lane * 4)
//   so that  address = (set offset in the upper bytes | packed byte << 8 | lane offset)  is still ONE byte-permute.
// Same arithmetic, same order of the floating-point updates per row (commit order): bit-identical to MODE 1.
#pragma once

#ifndef JW_WS_NB
#define JW_WS_NB 4                                   // builder warps
#endif
#define JW_WS_NS (JW_FUSED_THREADS / 32 - JW_WS_NB)  // streaming warps
#define JW_WS_SMEM (3 * 65536 + 4 * 96 * 4 + 64)     // tables | yq image of the slice | mbarriers

__device__ __forceinline__ void jw_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void jw_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool jw_mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: false when the sweep was abandoned (sticky abort flag / time-out), never a hang
__device__ __forceinline__ bool jw_mbar_wait(unsigned long long* bar, unsigned parity, int32_t* flags) {
    unsigned spins = 0; unsigned long long t0 = 0;
    while (!jw_mbar_test(bar, parity)) {
        if ((++spins & 1023u) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            int ab;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(ab) : "l"(flags + 2) : "memory");
            if (ab != 0) return false;
            if (now - t0 > 20000000000ull) { atomicExch(&flags[2], 1); return false; }
        }
    }
    return true;
}

// one streaming CTA = one row slice (vs) for the whole sweep; T = 1, 32-bit table entries
template <int T>
__device__ __forceinline__ void jw_stream_ws(const jw_fused_args& F, int* jw_smem, const int vs) {
    static_assert(T == 1, "the warp-specialised streaming role is built for one trait without missing calls");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Gs = F.Gs, R = Gs * 4;
    unsigned char* tab = reinterpret_cast<unsigned char*>(jw_smem);
    int* yqs = jw_smem + (3 * 65536) / 4;                                   // [R] (builder-private between its barriers)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(yqs + 4 * 96);   // full[2], empty[2]
    unsigned long long* full = bars; unsigned long long* empty = bars + 2;
    const int64_t nloc = F.nloc;
    float* const ycorr_l = F.ycorr + F.row_off;
    const int lag = F.lag;
    if (tid == 0) {
        jw_mbar_init(&full[0], 1); jw_mbar_init(&full[1], 1);
        jw_mbar_init(&empty[0], JW_WS_NS); jw_mbar_init(&empty[1], JW_WS_NS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t row0 = (int64_t)vs * R;

    if (warp >= JW_WS_NS) {
        // =============================== builder warps ===============================
        const int bt = tid - JW_WS_NS * 32;                               // 0 .. NB*32-1
        constexpr int NBT = JW_WS_NB * 32;
        constexpr int RPT = (JW_FUSED_MAX_GS * 4 + NBT - 1) / NBT;        // rows per builder thread (3)
        for (int k = 0; k < F.nblocks; ++k) {
            const int set = k & 1;
            if (k >= 2) { if (!jw_mbar_wait(&empty[set], (unsigned)(((k >> 1) - 1) & 1), F.flags)) return; }
            const int ap = k - 1 - lag;
            float v[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = bt + i * NBT;
                v[i] = (r < R && row0 + r < nloc) ? ycorr_l[row0 + r] : 0.0f;
            }
            int cnt = 0;
            if (ap >= 0) {
                // block ap's commits, in commit order, on this thread's rows (bytes from this CTA's tile of block ap)
                const int64_t s_ap = F.C.starts[ap];
                const int nch_ap = ((int)(F.C.starts[ap + 1] - s_ap) + 15) >> 4;
                const uint8_t* tile_ap = F.tiled + ((size_t)(F.chunk_off[ap] * F.n_vs + (int64_t)vs * nch_ap) * Gs) * 16;
                const bool okr = jw_rec_foreach<T>(F.P, F.P.blk_unit0[ap], F.P.blk_unit0[ap + 1],
                                                   [&](const int us_, const int nv, const jw_rec_reader<T>& RR) {
                    const int pbase = (int)(F.P.unit_start[us_] - s_ap);
                    unsigned bytes[JW_REC_BATCH][RPT]; float mus[JW_REC_BATCH];
#pragma unroll
                    for (int q = 0; q < JW_REC_BATCH; ++q) {
                        mus[q] = 0.0f;
#pragma unroll
                        for (int i = 0; i < RPT; ++i) bytes[q][i] = 0u;
                        if (q < nv) {
                            const int pm = pbase + RR.code(q);
                            mus[q] = F.C.means[s_ap + pm];
                            const uint8_t* col = tile_ap + ((size_t)(pm >> 4) * Gs << 4) + (pm & 15);
#pragma unroll
                            for (int i = 0; i < RPT; ++i) {
                                const int r = bt + i * NBT;
                                if (r < R) bytes[q][i] = col[(size_t)(r >> 2) << 4];
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < JW_REC_BATCH; ++q) {
                        if (q < nv) {
                            const float d = RR.d(q, 0);
#pragma unroll
                            for (int i = 0; i < RPT; ++i) {
                                const int r = bt + i * NBT;
                                const unsigned code = (bytes[q][i] >> ((r & 3) << 1)) & 3u;
                                const float xv = (code == 3u ? mus[q] : (float)code) - mus[q];
                                if (d != 0.0f && r < R && row0 + r < nloc) v[i] = fmaf(d, xv, v[i]);
                            }
                        }
                    }
                    cnt += nv;
                });
                if (!okr) { atomicExch(&F.flags[2], 1); return; }
            }
            // write the rows back, fixed-point image, sum of the image
            long long qs = 0; int ovf = 0;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = bt + i * NBT;
                if (r < R) {
                    const bool rv = row0 + r < nloc;
                    if (rv && cnt > 0) ycorr_l[row0 + r] = v[i];
                    const int q = rv ? jw_quantize(v[i], F.scale, &ovf) : 0;
                    yqs[r] = q; qs += q;
                }
            }
            if (ovf) atomicOr(&F.flags[0], 1);
            for (int o = 16; o > 0; o >>= 1) qs += __shfl_xor_sync(0xffffffffu, qs, o);
            if (lane == 0 && qs != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&F.sq_acc[k * T]), (unsigned long long)qs);
            asm volatile("bar.sync 1, %0;" :: "n"(NBT) : "memory");        // the image is complete (builder warps only)
            // table set `set`: entry e of group g = sum over its 4 individuals
            for (int item = bt; item < Gs * 16; item += NBT) {
                const int g = item % Gs, ehi = item / Gs;
                const unsigned c2 = ehi & 3, c3 = ehi >> 2;
                const int y0 = yqs[4 * g], y1 = yqs[4 * g + 1], y2 = yqs[4 * g + 2], y3 = yqs[4 * g + 3];
                const int Bv = jw_tabval(c2, y2, false) + jw_tabval(c3, y3, false);
                const int gb = g >> 5, l = g & 31;
                unsigned char* base = gb == 2 ? tab + 131072 + set * 128 + l * 4 : tab + set * 65536 + gb * 128 + l * 4;
#pragma unroll
                for (int elo = 0; elo < 16; ++elo) {
                    const int val = jw_tabval(elo & 3, y0, false) + jw_tabval(elo >> 2, y1, false) + Bv;
                    *reinterpret_cast<int*>(base + (ehi * 16 + elo) * 256) = val;
                }
            }
            asm volatile("bar.sync 1, %0;" :: "n"(NBT) : "memory");        // every builder thread's table stores are done
            if (bt == 0) jw_mbar_arrive(&full[set]);
            // pull the next panel's tile of this CTA into L2 (TMA bulk prefetch, no SM involvement after issue)
            if (warp == JW_WS_NS && k + 1 < F.nblocks && F.l2_prefetch) {
                const int nch1 = ((int)(F.C.starts[k + 2] - F.C.starts[k + 1]) + 15) >> 4;
                const uint8_t* t1 = F.tiled + ((size_t)(F.chunk_off[k + 1] * F.n_vs + (int64_t)vs * nch1) * Gs) * 16;
                const unsigned total = (unsigned)nch1 * Gs * 16;
                const unsigned per = ((total / 32) + 15) & ~15u;
                const unsigned off = per * lane;
                if (off < total) jw_prefetch_l2(t1 + off, min(per, total - off));
            }
        }
        return;
    }

    // =============================== streaming warps ===============================
    int64_t md_s = F.C.starts[0], md_e = F.C.starts[1], md_co = F.chunk_off[0];
    for (int k = 0; k < F.nblocks; ++k) {
        const int64_t s = md_s;
        const int b = (int)(md_e - md_s);
        const int64_t chunk_off_k = md_co;
        md_s = md_e;
        if (k + 1 < F.nblocks) { md_e = F.C.starts[k + 2]; md_co = F.chunk_off[k + 1]; }
        const int set = k & 1;
        const int nchunks = (b + 15) >> 4;
        const uint8_t* tile = F.tiled + ((size_t)(chunk_off_k * F.n_vs + (int64_t)vs * nchunks) * Gs) * 16;
        // set offset in the upper bytes, lane offset in the low byte: (packed byte << 8) is permuted in between
        const uint32_t lo0 = ((uint32_t)set << 16) | ((uint32_t)lane * 4u);
        const uint32_t lo1 = ((uint32_t)set << 16) | (128u + (uint32_t)lane * 4u);
        const uint32_t lo2 = (2u << 16) | ((uint32_t)set * 128u + (uint32_t)lane * 4u);
        if (!jw_mbar_wait(&full[set], (unsigned)((k >> 1) & 1), F.flags)) return;
        for (int mc = warp; mc < nchunks; mc += JW_WS_NS) {
            int acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = 0;
            uint4 dv[JW_FUSED_MAX_GS / 32];
#pragma unroll
            for (int gb = 0; gb < JW_FUSED_MAX_GS / 32; ++gb) {
                const int g = gb * 32 + lane;
                dv[gb] = make_uint4(0, 0, 0, 0);
                if (g < Gs) dv[gb] = __ldg(reinterpret_cast<const uint4*>(tile + ((size_t)(mc * Gs + g) << 4)));
            }
#pragma unroll
            for (int gb = 0; gb < JW_FUSED_MAX_GS / 32; ++gb) {
                const int g = gb * 32 + lane;
                if (g < Gs) {
                    const uint32_t laneoff = gb == 0 ? lo0 : (gb == 1 ? lo1 : lo2);
                    const uint32_t wds[4] = {dv[gb].x, dv[gb].y, dv[gb].z, dv[gb].w};
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        // result bytes: [0] lane offset, [1] packed byte q, [2] set / region, [3] 0
                        const uint32_t off = __byte_perm(wds[q >> 2], laneoff, 0x7604u | ((q & 3) << 4));
                        acc[q] += *reinterpret_cast<const int*>(tab + off);
                    }
                }
            }
            // transposed butterfly: 16 markers x 32 lanes -> marker (lane>>1)&15 on every lane
            long long vals[8];
            {
                const bool up = (lane & 16) != 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int keep = up ? acc[i + 8] : acc[i];
                    const int send = up ? acc[i] : acc[i + 8];
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
            const int jj = mc * 16 + q;
            if ((lane & 1) == 0 && jj < b && tot != 0)
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(&F.dq[s + jj]), "l"((unsigned long long)tot) : "memory");
        }
        // this warp's part of the panel is published by the warp itself (no CTA barrier): the lanes' reds are ordered
        // before lane 0's release by the warp barrier; the release is cumulative at gpu scope
        __syncwarp();
        if (lane == 0) {
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" :: "l"(&F.arrive[k]) : "memory");
            jw_mbar_arrive(&empty[set]);
        }
    }
}
