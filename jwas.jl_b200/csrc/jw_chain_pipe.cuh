// jw_chain_pipe.cuh -- the sequential Gibbs chain of the lagged exact schedule, pipelined over several
// chain CTAs of the persistent sweep kernel (engine 1, option chain_ctas >= 1).
//
// The chain is walked in UNITS of up to 1024 markers (one thread per marker; a panel of the exact
// schedule is 1..4 units).  Unit u is owned by chain CTA (u mod n_chain).  Everything that does not
// depend on earlier units of the same sweep -- state and constant loads, the block rhs, Gram-row
// prefetches, the write-back of alpha/beta/delta, the ordered active list -- runs off the critical
// path on the owning CTA while the previous units are still being decided by other CTAs.  What is
// sequential (BayesABC.jl:24-58 applied marker after marker) travels as COMMIT RECORDS:
//
//   one 64-bit word per (commit, trait):  [63:32] float bits of old-new alpha
//                                         [31:16] sweep tag   [15:0] position inside the unit
//   and one terminator word per unit (code 0x8000).
//
// A record is written with a single 64-bit store the moment the marker commits and is complete in
// itself, so no flag / fence / second round trip is needed: consumers poll the word until the tag
// is this sweep's.  Three consumers read the same records:
//   * later units of the same panel      (rhs correction with the panel's Gram block,  BayesABC.jl:169)
//   * the units of the next panel        (cross-Gram correction of the lagged schedule)
//   * the streaming CTAs, two panels on  (ycorr <- ycorr + (old-new) x_j,             BayesABC.jl:48)
// The additions happen in commit order everywhere, i.e. the same sums in the same order as the
// one-CTA chain (jw_chain_block) and the oracle's lagged schedule: results are bit-identical.
#pragma once
#include "jw_sweep_kernels.cuh"

#define JW_REC_END 0x8000u
#define JW_REC_STRIDE (JW_CHAIN_SB + 1)       // commits of one unit + terminator
#ifndef JW_REC_BATCH
#define JW_REC_BATCH 4                        // records peeked per poll (their Gram / genotype loads overlap)
#endif

struct jw_pipe_args {
    int n_chain;                              // chain CTAs (0 = one-CTA chain, jw_chain_block)
    int nunits;
    const int64_t* unit_start;                // nunits+1: first marker of every unit
    const int32_t* unit_blk;                  // nunits: block (panel) of the unit
    const int32_t* blk_unit0;                 // nblocks+1: first unit of every block
    unsigned long long* rec;                  // nunits * JW_REC_STRIDE * T words
    unsigned tag;                             // 1..65535, changes every sweep
    int32_t* act_cnt_unit;                    // nunits: entries of the unit's ordered active list
    int32_t* flags;                           // [2] sticky abort
    // optional back-off (ns) after an EMPTY poll of a record word (A/B knob: 0 / 300 / 1000 ns measured identical at
    // cfg2 in the sparse and in the dense regime, so polling pressure is not what bounds the chain)
    unsigned sleep_stream, sleep_chain;
};

__device__ __forceinline__ unsigned long long jw_ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void jw_st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long jw_rec_pack(float d, unsigned tag, unsigned code) {
    return ((unsigned long long)__float_as_uint(d) << 32) | ((unsigned long long)(tag & 0xffffu) << 16) | (code & 0xffffu);
}

// Reader of one unit's record stream.  issue() starts the loads of the next JW_REC_BATCH words at the
// cursor; resolve() looks at them: it returns how many commits (0..JW_REC_BATCH) have arrived, advances
// over them and sets `finished` once the terminator has been seen.  w[q][k] holds the words.  All lanes
// of a warp read the same addresses in the same instruction, so the result is warp-uniform; different
// warps may be at different cursors.
template <int T>
struct jw_rec_reader {
    const unsigned long long* base;
    unsigned tag;
    int e;
    bool finished;
    unsigned long long w[JW_REC_BATCH][T];

    __device__ __forceinline__ void open(const unsigned long long* rec, unsigned tag_, int unit) {
        base = rec + (size_t)unit * JW_REC_STRIDE * T;
        tag = tag_ & 0xffffu; e = 0; finished = false;
    }
    __device__ __forceinline__ void open(const jw_pipe_args& P, int unit) { open(P.rec, P.tag, unit); }
    __device__ __forceinline__ void issue() {
#pragma unroll
        for (int q = 0; q < JW_REC_BATCH; ++q)
#pragma unroll
            for (int k = 0; k < T; ++k)
                w[q][k] = (e + q < JW_REC_STRIDE) ? jw_ld_relaxed_u64(base + (size_t)(e + q) * T + k) : 0ull;
    }
    __device__ __forceinline__ int resolve() {
        int nv = 0;
#pragma unroll
        for (int q = 0; q < JW_REC_BATCH; ++q) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < T; ++k) ok = ok && (((unsigned)(w[q][k] >> 16) & 0xffffu) == tag);
            if (ok && nv == q && !finished) {
                if ((unsigned)w[q][0] & JW_REC_END) finished = true; else nv = q + 1;
            }
        }
        e += nv;
        return nv;
    }
    __device__ __forceinline__ int code(int q) const { return (int)((unsigned)w[q][0] & 0x7fffu); }
    __device__ __forceinline__ float d(int q, int k) const { return __uint_as_float((unsigned)(w[q][k] >> 32)); }
};

// spin bookkeeping shared by every record poller: returns false when the sweep must be abandoned
__device__ __forceinline__ bool jw_spin_ok(unsigned& spins, unsigned long long& t0, int32_t* flags) {
    if ((++spins & 255u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        int ab;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(ab) : "l"(flags + 2) : "memory");
        if (ab != 0) return false;
        if (now - t0 > 20000000000ull) { atomicExch(&flags[2], 1); return false; }
    }
    return true;
}

// Walks the commits of units [u0, u1) in commit order, handing them to fn(unit, nv, reader) in batches of
// up to JW_REC_BATCH (fn starts the batch's dependent loads together, then adds in order).  The first
// words of unit u+1 are requested before unit u is drained, so a run of finished units costs one memory
// round trip for the records plus one for what fn loads, not two per unit.  Returns false on abort.
// GENTLE pollers (the streaming CTAs' gather warps: ~150 of them, with slack) sleep between empty polls so
// that they neither take issue slots from the streaming warps nor hammer the L2 slice of the record being
// written; the chain CTAs (a handful, on the critical path) poll back to back.
template <int T, bool GENTLE = false, class BatchFn>
__device__ __forceinline__ bool jw_rec_foreach(const jw_pipe_args& P, const int u0, const int u1, BatchFn fn) {
    if (u0 >= u1) return true;
    unsigned spins = 0; unsigned long long t0 = 0;
    jw_rec_reader<T> cur, nxt;
    cur.open(P, u0); cur.issue();
    for (int u = u0; u < u1; ++u) {
        const bool more = u + 1 < u1;
        if (more) { nxt.open(P, u + 1); nxt.issue(); }
        while (true) {
            const int nv = cur.resolve();
            if (nv > 0) fn(u, nv, cur);
            if (cur.finished) break;
            if (nv == 0) {
                if (!jw_spin_ok(spins, t0, P.flags)) return false;
                if (GENTLE) __nanosleep(400);
                else if (P.sleep_stream) __nanosleep(P.sleep_stream);
            }
            cur.issue();
        }
        if (more) cur = nxt;
    }
    return true;
}

// Resumable walk over the commits of units [u0, u1) in commit order (the chain CTAs' view of the records):
// poll() returns  n > 0: the next n commits (of unit `unit`, words in R);  0: nothing new yet;  -1: all units done.
template <int T>
struct jw_rec_walker {
    const unsigned long long* rec;        // (no pointer to the argument struct: that would spill it to local memory)
    unsigned tag;
    int u, u1;
    bool loaded;
    jw_rec_reader<T> R;
    __device__ __forceinline__ void init(const jw_pipe_args& P_, int u0, int u1_) {
        rec = P_.rec; tag = P_.tag; u = u0; u1 = u1_; loaded = false;
        if (u < u1) R.open(rec, tag, u);
    }
    __device__ __forceinline__ int poll(int& unit) {
        while (true) {
            if (u >= u1) return -1;
            if (!loaded) R.issue();
            loaded = false;
            const int nv = R.resolve();
            if (nv > 0 || R.finished) {
                unit = u;
                if (R.finished) { u += 1; if (u < u1) R.open(rec, tag, u); }   // open() keeps the words of this batch
                if (nv > 0) return nv;
                continue;
            }
            return 0;
        }
    }
};

#define JW_UNIT_PG 16          // corrections fetched ahead of the block's rhs (per thread, in shared memory)
#define JW_UNIT_RING 16        // Gram rows of the predicted commits (markers already in the model) in flight / in shared memory
__host__ __device__ inline size_t jw_chain_unit_smem_bytes(int T) {
    return 512 + (size_t)T * JW_CHAIN_SB * 4 + (size_t)32 * JW_UNIT_PG * T * 4 +
           (size_t)JW_UNIT_PG * JW_CHAIN_SB * 4 + (size_t)JW_UNIT_RING * JW_CHAIN_SB * 4 + (size_t)JW_CHAIN_SB * 4;
}
// 4-byte asynchronous global -> shared copy (LDGSTS): the Gram-row ring is filled without holding registers
__device__ __forceinline__ void jw_cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void jw_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void jw_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// One unit of the chain.  B describes the unit's panel (s, b, Gram, cross-Gram towards the previous
// panel, rhs source); wait_fn() blocks until the panel's rhs partial sums are complete.
// Shared memory (caller-supplied, jw_chain_unit_smem_bytes): wmin[2][32] | cnt[32] | misc[32] | dc[T][1024] |
// pd[32 warps][PG][T] | pg[PG][1024] | ring[RING][1024] | list[1024].
//
// Everything that can be had before the block's rhs exists is fetched while the CTA would otherwise wait for
// it: state, constants, draws; and the cross-Gram / Gram values of every correction whose record has already been
// published (all of the previous panels, normally).  The markers that already carry an effect are the PREDICTED
// commits of the unit (a marker in the model changes its effect every time it is sampled): their Gram rows stream
// through a ring of JW_UNIT_RING rows in shared memory, filled by asynchronous copies (cp.async) that run RING
// commits ahead of the chain.  After the wait the critical path is: rhs partial sums (one L2 round trip) -> buffered
// corrections (shared memory) -> rounds (evaluation + one barrier + a shared-memory read each; a Gram round trip
// only when a marker ENTERS the model).
// Returns the number of commits, -1 when the sweep was aborted.
// Inlined into the persistent kernel on purpose: as a function of its own (measured) the chain runs ~2x slower --
// the call ABI and a second spill set under the kernel's 64-register cap land on the chain's critical path, and the
// chain, not the stream, bounds the exact schedule.  Rarely used paths called from here (the multi-GPU exchange
// reads) are therefore kept OUT of line so that they do not disturb this function's register allocation.
template <int METHOD, int T, bool MULTI, class WaitFn>
__device__ __forceinline__ int jw_chain_unit(const jw_chain_args& A, const jw_pipe_args& P, const jw_chain_blk& B,
                                             const int u, WaitFn wait_fn, unsigned char* smem_base,
                                             unsigned long long* ct /* 5 phase timers or nullptr */) {
    int (*s_wmin)[32] = reinterpret_cast<int (*)[32]>(smem_base);
    int* s_cnt = reinterpret_cast<int*>(smem_base + 256);
    int* s_misc = reinterpret_cast<int*>(smem_base + 384);                    // [0] cached rows, [1..ROWS] their positions
    float* s_dc = reinterpret_cast<float*>(smem_base + 512);                  // [T][JW_CHAIN_SB]
    float* s_pd = s_dc + T * JW_CHAIN_SB;                                     // [32][PG][T]
    float* s_pg = s_pd + 32 * JW_UNIT_PG * T;                                 // [PG][JW_CHAIN_SB]
    float* s_ring = s_pg + JW_UNIT_PG * JW_CHAIN_SB;                          // [RING][JW_CHAIN_SB]
    int* s_list = reinterpret_cast<int*>(s_ring + JW_UNIT_RING * JW_CHAIN_SB); // [JW_CHAIN_SB] ordered positions of the predicted commits
    (void)s_misc;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = (int)(blockDim.x >> 5);
    const int64_t p = A.p;
    const int kb = P.unit_blk[u];
    const int64_t s = B.s; const int b = B.b;
    const int64_t us = P.unit_start[u];
    const int m0 = (int)(us - s);                         // unit's first position inside the panel
    const int ub = min(JW_CHAIN_SB, b - m0);
    const int m = m0 + tid;
    const bool valid = tid < ub;
    const int64_t j = us + (valid ? tid : 0);
    const float* G = A.gram + B.gram_off;
    unsigned long long* myrec = P.rec + (size_t)u * JW_REC_STRIDE * T;
    unsigned long long ctm = 0;
    const bool ctimed = (ct != nullptr) && tid == 0;
    if (ctimed) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ctm));
#define JW_CT(i) do { if (ctimed) { unsigned long long n__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n__)); ct[i] += n__ - ctm; ctm = n__; } } while (0)

    // the inputs of this CTA's NEXT unit are needed n_chain units from now: pull them towards L2 in bulk
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
                case 12: case 13:
                    if (A.draws_u) { base = (lane == 12 ? A.draws_u : A.draws_z) + kk * p + ps; esz = 8; } break;
                default: break;
            }
            if (base) {
                const unsigned long long a0 = (unsigned long long)base & ~15ull;
                const unsigned bytes = (unsigned)((((unsigned long long)base + (unsigned long long)pb * esz + 15ull) & ~15ull) - a0);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
            }
        }
    }

    // ---- state at unit entry, constants, draws (nothing here depends on other units) ----
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
    E.load_draws(A, j, 0);
    bool row_requested = false;
    bool nz = false;
    if (valid) {
#pragma unroll
        for (int k = 0; k < T; ++k) nz = nz || (a_cur[k] != 0.0f);
        if (nz) {            // certain to need its Gram row: one bulk prefetch towards L2 (later units read it too)
            row_requested = true;
            const float* row = G + (int64_t)m * b;
            const unsigned long long a0 = (unsigned long long)row & ~15ull;
            const unsigned bytes = (unsigned)((((unsigned long long)(row + b) + 15ull) & ~15ull) - a0);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
        }
    }
    // The PREDICTED commits of the unit = the markers already in the model (a marker in the model changes its effect
    // every time it is sampled), and their Gram rows (this unit's columns only: thread tid copies and later reads
    // column tid): the rows start streaming into the ring while the CTA waits for the rhs.  (Predicting the markers about
    // to ENTER the model as well, from one extra evaluation after the rhs, was measured and is not worth its barrier:
    // 50.5 vs 47.3 ms per sweep with pi fixed at 0.95.)
    int nlist = 0;
    auto build_list = [&](const bool flag) {
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        const int c = (lane < nw) ? s_cnt[lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int vv = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += vv; }
        const int warp_off = __shfl_sync(0xffffffffu, incl - c, warp);
        nlist = __shfl_sync(0xffffffffu, incl, 31);
        if (flag) s_list[warp_off + __popc(bal & ((1u << lane) - 1u))] = tid;
        __syncthreads();
    };
    auto ring_issue = [&](const int i) {      // row of predicted commit i -> slot i % RING (an empty group when there is none)
        if (i < nlist && valid) jw_cp_async4(s_ring + (i % JW_UNIT_RING) * JW_CHAIN_SB + tid, G + (int64_t)(m0 + s_list[i]) * b + m);
        jw_cp_async_commit();
    };
    build_list(nz);
    for (int i = 0; i < JW_UNIT_RING; ++i) ring_issue(i);
    int li = 0;                                // next predicted commit (uniform across the CTA)

    // corrections whose records are already there: fetch their (cross-)Gram values now, add them after the rhs
    // corrections come from the units of the `lag` previous panels (oldest first) and the earlier units of this one
    const int u_own = P.blk_unit0[kb];
    const int u_l1 = (kb > 0 && B.xgram != nullptr) ? P.blk_unit0[kb - 1] : u_own;
    const int u_lo = (kb > 1 && B.xgram2 != nullptr) ? P.blk_unit0[kb - 2] : u_l1;
    jw_rec_walker<T> W;
    W.init(P, u_lo, u);
    int pgn = 0, wst = 0;
    float* s_pd_w = s_pd + warp * (JW_UNIT_PG * T);
    auto corr_row = [&](const int us_, const int code) -> const float* {
        // units are cut every JW_CHAIN_SB markers inside a panel: no table look-up on this path
        const int ubase = us_ >= u_own ? u_own : (us_ >= u_l1 ? u_l1 : u_lo);
        const float* M = us_ >= u_own ? G : (us_ >= u_l1 ? B.xgram : B.xgram2);
        return M + (int64_t)((us_ - ubase) * JW_CHAIN_SB + code) * b + m;
    };
    while (pgn + JW_REC_BATCH <= JW_UNIT_PG) {
        int us_ = 0;
        wst = W.poll(us_);
        if (wst <= 0) break;
#pragma unroll
        for (int q = 0; q < JW_REC_BATCH; ++q) {
            if (q < wst) {
                s_pg[(pgn + q) * JW_CHAIN_SB + tid] = valid ? *corr_row(us_, W.R.code(q)) : 0.0f;
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < T; ++k) s_pd_w[(pgn + q) * T + k] = W.R.d(q, k);
                }
            }
        }
        pgn += wst;
    }
    JW_CT(0);
    if (!wait_fn()) return -1;
    JW_CT(1);

    // ---- rhs of this marker from the streamed partial sums ----
    const int mc = valid ? m : m0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        long long dq, mq, sqk;
        if (MULTI) {
            dq = 0; mq = 0; sqk = 0;
            ok = jw_ll_rhs(B, T, k, mc, A.mq != nullptr, dq, mq, sqk) && ok;
        } else {
            dq = __ldcg(&A.dq[k * p + j]); mq = A.mq ? __ldcg(&A.mq[k * p + j]) : 0ll;
            sqk = __ldcg(&B.sq[k]);
        }
        r[k] = ((double)dq - mu * (double)(sqk - mq)) * A.invscale;
    }

    // ---- corrections in commit order: the buffered ones, then whatever is still being decided ----
    __syncwarp();
    for (int e = 0; e < pgn; ++e) {
        const float g = s_pg[e * JW_CHAIN_SB + tid];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            const float d = s_pd_w[e * T + k];
            if (d != 0.0f) r[k] += (double)d * (double)g;
        }
    }
    {
        unsigned spins = 0; unsigned long long t0 = 0;
        while (wst >= 0 && ok) {
            int us_ = 0;
            wst = W.poll(us_);
            if (wst < 0) break;
            if (wst == 0) {
                if (!jw_spin_ok(spins, t0, P.flags)) { ok = false; break; }
                if (P.sleep_chain) __nanosleep(P.sleep_chain);
                continue;
            }
            float g[JW_REC_BATCH];
#pragma unroll
            for (int q = 0; q < JW_REC_BATCH; ++q) g[q] = (q < wst && valid) ? *corr_row(us_, W.R.code(q)) : 0.0f;
#pragma unroll
            for (int q = 0; q < JW_REC_BATCH; ++q) {
                if (q < wst) {
#pragma unroll
                    for (int k = 0; k < T; ++k) {
                        const float d = W.R.d(q, k);
                        if (d != 0.0f) r[k] += (double)d * (double)g[q];
                    }
                }
            }
        }
    }
    if (__syncthreads_or(ok ? 0 : 1)) return -1;          // a poller gave up: the whole CTA leaves together
    JW_CT(2);

    // ---- speculative rounds (one barrier each), every commit published at once ----
    // (Measured and dropped: evaluating only a 32/64/128-marker window per round in units with many markers in the model
    // -- 46.1 vs 46.1 ms per sweep with pi fixed at 0.95.  A round costs ~1 us because all 32 warps of the CTA run the
    // loop's ~150 instructions of bookkeeping every round, whatever they evaluate; see DESIGN.md, dense regime.)
    unsigned long long my_active = 0, my_rounds = 0;
    int parity = 0, ncommit = 0, pos = 0;
    while (true) {
        const bool pending = valid && tid >= pos;
        float newA[T], newB[T]; int newD[T];
        bool active = false;
        if (pending) {
            active = E.eval(A, r, a_cur, b_cur, d_cur, newA, newB, newD);
            if (active) {
#pragma unroll
                for (int k = 0; k < T; ++k) s_dc[k * JW_CHAIN_SB + tid] = a_cur[k] - newA[k];
                if (!row_requested) {
                    row_requested = true;
                    const float* row = G + (int64_t)m * b;
                    const unsigned long long a0 = (unsigned long long)row & ~15ull;
                    const unsigned bytes = (unsigned)((((unsigned long long)(row + b) + 15ull) & ~15ull) - a0);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
                }
            }
        }
        const int key = (pending && active) ? tid : 0x7fffffff;
        const int wmin = __reduce_min_sync(0xffffffffu, key);
        if (lane == 0) s_wmin[parity][warp] = wmin;
        __syncthreads();
        const int v = (lane < nw) ? s_wmin[parity][lane] : 0x7fffffff;
        const int first = __reduce_min_sync(0xffffffffu, v);
        parity ^= 1;
        my_rounds += (tid == 0);
        if (pending && tid == first) {
            // the committing marker publishes its own record before anything else happens
#pragma unroll
            for (int k = 0; k < T; ++k)
                jw_st_relaxed_u64(myrec + (size_t)ncommit * T + k, jw_rec_pack(a_cur[k] - newA[k], P.tag, (unsigned)first));
            my_active += 1;
        }
        if (pending && tid <= first) {
#pragma unroll
            for (int k = 0; k < T; ++k) { a_cur[k] = newA[k]; b_cur[k] = newB[k]; d_cur[k] = newD[k]; }
        }
        if (first == 0x7fffffff) break;
        const int fg = m0 + first;                       // committed marker's position inside the panel
        // the committed marker's Gram row: from the ring when it was predicted (the usual case), else from L2.
        // Every predicted commit at or before `first` retires its ring slot: wait for the oldest copy group, take the
        // value, refill the slot with the row RING predictions ahead (keeps exactly RING groups in flight).
        float g = 0.0f; bool hit = false;
        while (li < nlist && s_list[li] <= first) {
            jw_cp_async_wait<JW_UNIT_RING - 1>();
            if (s_list[li] == first) { hit = true; g = s_ring[(li % JW_UNIT_RING) * JW_CHAIN_SB + tid]; }
            ring_issue(li + JW_UNIT_RING);
            li += 1;
        }
        if (valid && tid > first) {
            if (!hit) g = G[(int64_t)fg * b + m];
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const float d = s_dc[k * JW_CHAIN_SB + first];
                if (d != 0.0f) r[k] += (double)d * (double)g;
            }
        }
        if (B.xgram_next != nullptr && tid == 0) {
            // the next panel's units will need this marker's cross-Gram row: start moving it to L2
            const float* row = B.xgram_next + (int64_t)fg * B.b_next;
            const unsigned long long a0 = (unsigned long long)row & ~15ull;
            const unsigned bytes = (unsigned)((((unsigned long long)(row + B.b_next) + 15ull) & ~15ull) - a0);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
        }
        ncommit += 1;
        pos = first + 1;
        if (pos >= ub) break;
    }
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < T; ++k) jw_st_relaxed_u64(myrec + (size_t)ncommit * T + k, jw_rec_pack(0.0f, P.tag, JW_REC_END));
    }
    JW_CT(3);

    // ---- off the critical path: publish state, net delta-alpha and the ordered active list ----
    bool any = false;
    if (valid) {
#pragma unroll
        for (int k = 0; k < T; ++k) {
            A.alpha[k * p + j] = a_cur[k];
            if (METHOD != 1) A.beta[k * p + j] = b_cur[k];
            A.delta[k * p + j] = d_cur[k];
            const float d = a_entry[k] - a_cur[k];
            A.dalpha[k * p + j] = d;
            any = any || (d != 0.0f);
        }
    }
    int act_total = 0;
    if (__syncthreads_or(any ? 1 : 0)) {
        const unsigned bal = __ballot_sync(0xffffffffu, any);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        const int c = (lane < nw) ? s_cnt[lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int vv = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += vv; }
        const int warp_off = __shfl_sync(0xffffffffu, incl - c, warp);
        act_total = __shfl_sync(0xffffffffu, incl, 31);
        if (any) B.act_idx[warp_off + __popc(bal & ((1u << lane) - 1u))] = (int32_t)j;
    }
    if (tid == 0) P.act_cnt_unit[u] = act_total;
    jw_cp_async_wait<0>();    // no copy of this unit may land in the ring after the next unit starts filling it
    __syncthreads();          // shared scratch is reused by this CTA's next unit
    JW_CT(4);
#undef JW_CT
    if (A.counters) {
        if (my_active) atomicAdd(&A.counters[0], my_active);
        if (my_rounds) atomicAdd(&A.counters[1], my_rounds);
    }
    return ncommit;
}
