// jw_fused_sweep.cuh -- "engine 1": the whole marker sweep as ONE persistent cooperative kernel.
//
// Every CTA owns a slice of individuals (rows of M) for the whole sweep.  Per block of markers:
//   1. apply the previous block's delta-alpha to the slice of ycorr (fused axpy, BayesABC.jl:48/184),
//      re-quantise it and rebuild a shared-memory lookup table: for each group of 4 individuals
//      (= one packed byte) the 256 possible partial dot products  sum_k value(code_k) * yq_k;
//   2. stream the block's packed genotypes (re-tiled so that a warp reads 512 contiguous bytes:
//      lane = byte-group, 16 markers per 128-bit load) and turn every byte into ONE table lookup
//      + ONE integer add -- 4 genotypes per lookup instead of 3+ instructions per genotype;
//   3. reduce across lanes and add the exact int64 partial rhs into global memory with
//      red.add.u64 (integer atomics commute: any CTA order gives the same bits);
//   4. the chain: lag 0 -- CTA 0 waits for all slices (release/acquire counter), runs the in-block Gibbs
//      chain (jw_chain_block, shared with engine 0) and publishes the block's ordered active list;
//      lag 1 -- dedicated chain CTA(s) do that while the other CTAs already stream the next block: one
//      chain CTA (MODE 0), or several that walk units of 1024 markers and hand each other 64-bit commit
//      records (MODE 1 / 2, jw_chain_pipe.cuh); the streaming CTAs replay the same records for the axpy.
// No host round trip, no kernel boundary, ycorr never leaves L2 for the duration of the sweep.
// Spin loops carry a time-out that raises a sticky abort flag instead of hanging the device.
#pragma once
#include "jw_common.cuh"
#include "jw_sweep_kernels.cuh"
#include "jw_chain_pipe.cuh"
#include <cooperative_groups.h>

#ifdef JW_NEXT
#define JW_NEXT_META
#define JW_NEXT_L1PF
#define JW_NEXT_RED
#endif
#define JW_FUSED_THREADS 1024
// multi-GPU exchange ring: a sender may run lag+3 blocks ahead of the slowest reader's chain
#define JW_X_RING 8
#define JW_FUSED_MAX_GS 96          // 3 lookups per lane and marker keep the int32 partial < 2^31

struct jw_fused_state {
    uint8_t* d_tiled = nullptr;
    int64_t* d_chunk_off = nullptr;  // nblocks+1, in 16-marker chunks
    int32_t* d_chunk_block = nullptr;
    int* d_arrive = nullptr;         // nblocks
    int* d_done = nullptr;           // 1
    long long* d_sq_acc = nullptr;   // nblocks * T
    int32_t* d_act_cnt_blk = nullptr;// nblocks
    // pipelined chain (option chain_ctas): units of <= 1024 markers, commit records
    int n_chain = 0, nunits = 0;
    int64_t* d_unit_start = nullptr; int32_t* d_unit_blk = nullptr; int32_t* d_blk_unit0 = nullptr;
    unsigned long long* d_rec = nullptr; size_t rec_bytes = 0;
    int32_t* d_act_cnt_unit = nullptr;
    unsigned rec_tag = 0;
    std::vector<int64_t> unit_start;
    int Gs = 0, TS = 0, n_vs = 0, n_cta = 0, W = 1, list_cap = 0, two_lists = 0;
    int64_t total_chunks = 0;
    size_t smem = 0, smem_pipe = 0;
    bool legacy_ok = true;           // the one-chain-CTA layout (with its commit lists) fits in shared memory
    bool ready = false;
};

struct jw_fused_args {
    jw_chain_args C;
    const uint8_t* tiled;
    const int64_t* chunk_off;
    const uint8_t* packed; int64_t stride_d;
    int Gs, TS, n_vs, nblocks, list_cap, lag;
    int uniform_b;               // > 0: every block has this many markers (the last one possibly fewer)
    int gather;                  // 1 = a gather warp replays the records under the stream (else: in line, before the tables)
    const float* gramx; const int64_t* gramx_off;        // cross-Gram X_{k-1}'X_k
    const float* gramx2; const int64_t* gramx2_off;      // cross-Gram X_{k-2}'X_k (lag 2), else NULL
    int timers, two_lists, l2_prefetch;
    int arrive_mult;             // arrivals per streaming CTA and panel (1; the warp-specialised role: one per streaming warp)
    // multi-GPU (rows sharded over `world` GPUs of one node, one process each): this rank streams the
    // byte-group slices [vs0, vs1) of the rows it stores; a communication CTA pushes the block's exact int64
    // partial rhs into every peer's exchange slots over NVLink (IPC-mapped peer memory) as self-validating
    // 16-byte words (value + tag): no flag, no fence; the chain's own rhs read is the wait.
    int world, rank, vs0, vs1;
    uint4* const* peer_slots;           // world pointers: each rank's exchange buffer as seen from this GPU
    const uint4* my_slots;              // this rank's own buffer (local address)
    int64_t slot_stride; int slot_b; int64_t ring_stride; unsigned tag_base;   // strides in 16-byte words
    int64_t row_off, nloc;              // this rank stores rows [row_off, row_off + nloc) of the n individuals
    float* ycorr; float scale;
    int* arrive; int* done; long long* sq_acc; int32_t* act_cnt_blk; int32_t* act_idx_all;
    int32_t* flags;              // [0] overflow, [2] abort
    long long* dq; long long* mq;
    jw_pipe_args P;
};

// ---- re-tiling: marker-major .jgb2 image -> [block][row slice][16-marker chunk][byte group][16] ----
__global__ void __launch_bounds__(256)
jw_k_tile(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t nbytes,
          const int64_t* __restrict__ starts, const int64_t* __restrict__ chunk_off,
          const int32_t* __restrict__ chunk_block, int64_t total_chunks, int Gs, int n_vs,
          uint8_t* __restrict__ tiled) {
    int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (gc, vs, g), g fastest
    int64_t total = total_chunks * n_vs * Gs;
    if (unit >= total) return;
    int g = (int)(unit % Gs);
    int64_t r = unit / Gs;
    int vs = (int)(r % n_vs);
    int64_t gc = r / n_vs;
    int k = chunk_block[gc];
    int64_t mc = gc - chunk_off[k];
    int64_t nchunks = chunk_off[k + 1] - chunk_off[k];
    int64_t s = starts[k], e = starts[k + 1];
    int64_t byteidx = (int64_t)vs * Gs + g;
    uint32_t w[4] = {0, 0, 0, 0};
    if (byteidx < nbytes) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            int64_t j = s + mc * 16 + q;
            if (j < e) w[q >> 2] |= (uint32_t)packed[j * stride_d + byteidx] << (8 * (q & 3));
        }
    }
    size_t off = ((size_t)(chunk_off[k] * n_vs + (int64_t)vs * nchunks) * Gs + (size_t)(mc * Gs + g)) * 16;
    *reinterpret_cast<uint4*>(tiled + off) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ int jw_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void jw_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// asynchronous HBM -> L2 prefetch of a contiguous region (TMA bulk prefetch, no SM involvement
// after issue); bytes must be a multiple of 16
__device__ __forceinline__ void jw_prefetch_l2(const void* ptr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(ptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long jw_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// thread-0 spin until *p >= target; false on time-out/abort
__device__ __forceinline__ bool jw_spin_ge(const int* p, int target, int32_t* flags) {
    unsigned long long t0 = jw_globaltimer();
    unsigned it = 0;
    while (jw_ld_acquire(p) < target) {
        if ((++it & 1023u) == 0) {
            if (jw_ld_acquire(&flags[2]) != 0) return false;
            if (jw_globaltimer() - t0 > 20000000000ull) { atomicExch(&flags[2], 1); return false; }
        }
    }
    return true;
}

// value a code contributes: component 0 = additive value (0,1,2,0); MISS component = indicator of 3
__device__ __forceinline__ int jw_tabval(unsigned c, int y, bool miss_comp) {
    if (miss_comp) return c == 3u ? y : 0;
    return c == 3u ? 0 : (int)c * y;
}

// Lookup-table layout.  Entry e of byte-group g (g = gb*32 + l) lives at byte offset
//   e*256 + sub(gb) + l*4*W (+4 for the second component):
// every table row is 256 bytes, so (byte << 8) | lane_offset is ONE byte-permute of the packed word,
// and the 32 lanes of a warp always hit 32 different banks (conflict-free by construction).
//   W=1: sub = {0, 128, 65536}   (128 KB)        W=2: sub = {0, 65536, 131072}   (192 KB)
template <int W>
__host__ __device__ __forceinline__ constexpr int jw_tab_sub(int gb) {
    return W == 1 ? (gb == 0 ? 0 : (gb == 1 ? 128 : 65536)) : gb * 65536;
}
#define JW_TAB_BYTES(W_) ((W_) == 1 ? 2 * 65536 : 3 * 65536)

#include "jw_fused_ws.cuh"

// communication CTA (rows sharded over several GPUs): the block's exact int64 partial rhs of this GPU -> every rank's
// exchange slots, as self-validating 16-byte words.  Value e of the slot: dq[kk][mm] | mq[kk][mm] (only when calls
// are missing) | sq[kk]; only the entries the chain reads are sent (mm < b).  Not inlined (see jw_chain_unit).
template <int T>
__device__ __noinline__ void jw_comm_push(const long long* dq, const long long* mq, const long long* sq_k, uint4* const* peer_slots,
                                          const int world, const int64_t slot0, const int slot_b, const unsigned tag,
                                          const int64_t p, const int64_t s, const int b) {
    const int tid = threadIdx.x;
    const int nmq = mq ? 2 : 1;
    for (int e = tid; e < nmq * T * b + T; e += JW_FUSED_THREADS) {
        long long v; int64_t w;
        if (e < nmq * T * b) {
            const int part = e / (T * b), e2 = e - part * T * b, kk = e2 / b, mm = e2 - kk * b;
            v = part == 0 ? __ldcg(&dq[(int64_t)kk * p + s + mm]) : __ldcg(&mq[(int64_t)kk * p + s + mm]);
            w = (int64_t)(part * T + kk) * slot_b + mm;
        } else {
            const int kk = e - nmq * T * b;
            v = __ldcg(&sq_k[kk]);
            w = (int64_t)2 * T * slot_b + kk;
        }
        for (int rk = 0; rk < world; ++rk) jw_ll_store(peer_slots[rk] + slot0 + w, v, tag);
    }
}

// ... and the way back: the block's partial rhs of EVERY rank (pushed into this GPU's slots by the ranks' communication
// CTAs) summed -- exact int64, any order -- and written over this GPU's own partial sums, so that the chain reads
// complete sums from dq / mq / sq exactly as on one GPU (no exchange code, and no extra registers, inside the chain).
// Returns false when the sweep was abandoned.
template <int T>
__device__ __noinline__ bool jw_comm_pull(long long* dq, long long* mq, long long* sq_k, const uint4* my_slots, const int world,
                                          const int64_t ring_off, const int64_t slot_stride, const int slot_b, const unsigned tag,
                                          const int64_t p, const int64_t s, const int b, int32_t* flags) {
    const int tid = threadIdx.x;
    const int nmq = mq ? 2 : 1;
    bool ok = true;
    for (int e = tid; e < nmq * T * b + T && ok; e += JW_FUSED_THREADS) {
        long long* dst; int64_t w;
        if (e < nmq * T * b) {
            const int part = e / (T * b), e2 = e - part * T * b, kk = e2 / b, mm = e2 - kk * b;
            dst = (part == 0 ? dq : mq) + (int64_t)kk * p + s + mm;
            w = (int64_t)(part * T + kk) * slot_b + mm;
        } else {
            const int kk = e - nmq * T * b;
            dst = sq_k + kk;
            w = (int64_t)2 * T * slot_b + kk;
        }
        long long tot = 0;
        ok = jw_ll_sum(my_slots + ring_off + w, slot_stride, world, tag, flags, tot);
        if (ok) __stcg(dst, tot);
    }
    return ok;
}

// MODE 0: one chain CTA (lag 1) or CTA 0 streams and chains (lag 0), jw_chain_block
//      1: pipelined chain (jw_chain_pipe.cuh), streaming CTAs replay the commit records in line
//      2: pipelined chain, one gather warp per streaming CTA replays them under the stream (one slice per CTA)
// The modes are compile-time: code of an unused role costs the streaming loop registers (measured: -13 %).
// MULTI (rows sharded over several GPUs) is compile-time as well: the single-GPU instantiations contain no exchange code.
template <int METHOD, int T, int W, int MODE, bool MULTI>
__global__ void __launch_bounds__(JW_FUSED_THREADS, 1)
jw_k_fused(jw_fused_args F) {
    extern __shared__ __align__(16) int jw_smem[];
    constexpr bool MISS = (T == 1 && W == 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = JW_FUSED_THREADS / 32;
    const int Gs = F.Gs, TS = F.TS, R = Gs * 4;
    unsigned char* tab = reinterpret_cast<unsigned char*>(jw_smem);
    int* yqs = jw_smem + JW_TAB_BYTES(W) / 4;    // [T][R]
    const int TRp = (T * R + 3) & ~3;
    float* s_y = reinterpret_cast<float*>(yqs + TRp);       // [T][R] this CTA's rows of ycorr (gather mode)
    float* s_yn = s_y + TRp;                                 // [T][R] the same rows after the next replay
    __shared__ unsigned s_tmax;                              // (timers) last streaming warp's finish time in the block
    __shared__ int s_gcnt[2];                                // commits replayed into s_yn by the gather warp, by block parity
    (void)TS;
    __shared__ long long s_red[32 * JW_MAX_TRAITS];
    __shared__ int s_ok;
    const int64_t n = F.C.n, p = F.C.p;
    // rows are LOCAL (0 .. nloc) in everything that touches the genotype image; ycorr holds all n individuals
    const int64_t nloc = F.nloc;
    float* const ycorr_l = F.ycorr + F.row_off;
    // lag = 1: the last CTA only runs chains, the others only stream, so that the chain of block k
    // overlaps the streaming of block k+1 (which needs the updates of blocks <= k-1 only)
    const int lag = F.lag;
    constexpr bool multi = MULTI;                        // host: MULTI == (world > 1); needs lag >= 1
    // pipelined chain: the last n_chain CTAs walk the chain unit by unit (jw_chain_pipe.cuh)
    constexpr bool pipe = MODE >= 1;
    const int n_chain = pipe ? F.P.n_chain : 1;
    const int n_stream = lag ? (int)gridDim.x - n_chain - (multi ? 1 : 0) : (int)gridDim.x;
    const bool is_chain_cta = lag ? (blockIdx.x >= gridDim.x - n_chain) : (blockIdx.x == 0);
    const bool is_comm_cta = multi && (blockIdx.x == gridDim.x - n_chain - 1);
    const bool is_stream_cta = lag ? (blockIdx.x < (unsigned)n_stream) : true;
    const int n_vs_local = F.vs1 - F.vs0;
    const bool single = n_vs_local <= n_stream;
    // phase timers (ns): [0] wait for previous chain, [1] axpy+quantise+tables, [2] stream,
    // [3] wait for all slices, [4] chain; CTA 0 -> counters[32..36], CTA 1 -> counters[40..44]
    unsigned long long ph[5] = {0, 0, 0, 0, 0};
#ifdef JW_TIMERS          // phase timers are a build option (tools/phase_probe.py): their mere presence costs ~10 %
    const bool timed = (tid == 0) && (blockIdx.x <= 1 || is_chain_cta) && (F.C.counters != nullptr) && F.timers;
    const bool timers_on = F.timers != 0;
#else
    constexpr bool timed = false;
    constexpr bool timers_on = false;
#endif
    unsigned long long tm = timed ? jw_globaltimer() : 0;
#define JW_PHASE(i) do { if (timed) { unsigned long long now__ = jw_globaltimer(); ph[i] += now__ - tm; tm = now__; } } while (0)
    int prev_commits = 0;                  // chain CTA, lagged schedule: commits of the previous block (smem list)
    long long sq_keep[T];                  // thread 0: this CTA's sum of yq over its slice(s)
#pragma unroll
    for (int kk = 0; kk < T; ++kk) sq_keep[kk] = 0;

    // gather mode (pipelined chain, one slice per CTA): the last warp does not stream; while the others
    // stream block k it replays the commit records of block k-1 against this CTA's rows of ycorr (kept
    // in shared memory for the whole sweep), so that block k+1 starts from finished values and the
    // record / genotype-byte round trips stay off the streaming CTAs' critical path
    const bool gather_mode = (MODE == 2) && is_stream_cta && single;
    const int nws = gather_mode ? nwarps - 1 : nwarps;                // streaming warps
    if (tid == 0) { s_gcnt[0] = 0; s_gcnt[1] = 0; s_tmax = 0; }
    bool gather_failed = false;

    if constexpr (MODE >= 1) { if (is_chain_cta) {
        // ---- pipelined chain: this CTA owns units cidx, cidx + n_chain, ... ----
        const int cidx = (int)blockIdx.x - ((int)gridDim.x - n_chain);
        unsigned long long ct[5] = {0, 0, 0, 0, 0};
        const bool ctimed = timed && cidx == 0;
        int my_units = 0;
        for (int u = cidx; u < F.P.nunits; u += n_chain, ++my_units) {
            const int k = F.P.unit_blk[u];
            jw_chain_blk B;
            B.s = F.C.starts[k]; B.b = (int)(F.C.starts[k + 1] - B.s); B.gram_off = F.C.gram_off[k];
            B.xgram = nullptr; B.xgram2 = nullptr; B.xlist = nullptr; B.xcount = nullptr; B.xstart = 0; B.xgram_next = nullptr; B.b_next = 0;
            B.xcount_smem = -1;
            if (k > 0) { B.xgram = F.gramx + F.gramx_off[k]; B.xstart = F.C.starts[k - 1]; }
            if (k > 1 && F.gramx2 != nullptr) B.xgram2 = F.gramx2 + F.gramx2_off[k];
            if (k + 1 < F.nblocks) {
                B.xgram_next = F.gramx + F.gramx_off[k + 1];
                B.b_next = (int)(F.C.starts[k + 2] - F.C.starts[k + 1]);
            }
            B.prefetch_s = 0; B.prefetch_b = 0;
            if (u + n_chain < F.P.nunits) {
                B.prefetch_s = F.P.unit_start[u + n_chain];
                B.prefetch_b = (int)(F.P.unit_start[u + n_chain + 1] - B.prefetch_s);
            }
            B.xslots = nullptr; B.xworld = 1; B.slot_stride = 0; B.slot_b = 0; B.xtag = 0; B.xflags = F.flags;
            B.sq = F.sq_acc + k * T;
            B.act_idx = F.act_idx_all + F.P.unit_start[u];
            B.act_cnt = nullptr; B.write_active_list = 1;
            auto wait_rhs = [&]() -> bool {
                // one GPU: every streaming CTA (or warp) has arrived; several: + the communication CTA, which has
                // then replaced this GPU's partial sums in dq / mq / sq by the sums over all ranks
                if (tid == 0) s_ok = jw_spin_ge(&F.arrive[k], n_stream * F.arrive_mult + (multi ? 1 : 0), F.flags) ? 1 : 0;
                __syncthreads();
                return s_ok != 0;
            };
            const int nc = jw_chain_unit<METHOD, T, false>(F.C, F.P, B, u, wait_rhs,
                                                    reinterpret_cast<unsigned char*>(jw_smem), ctimed ? ct : nullptr);
            if (nc < 0) return;
        }
        if (ctimed) {
            // [56..60] preload | wait for the panel's rhs | rhs + record corrections | rounds | epilogue; [61] units
            for (int i = 0; i < 5; ++i) F.C.counters[56 + i] = ct[i];
            F.C.counters[61] = (unsigned long long)my_units;
        }
        return;
    } }

    if constexpr (MODE == 3) { if (is_stream_cta) {
        // warp-specialised streaming role (jw_fused_ws.cuh): builder warps and streaming warps, two table sets
        if ((int)blockIdx.x < n_vs_local) jw_stream_ws<T>(F, jw_smem, F.vs0 + (int)blockIdx.x);
        return;
    } }

    // block metadata is fetched one iteration ahead: a dependent global load costs ~1 us inside this kernel
    int64_t md_s = F.C.starts[0], md_e = F.C.starts[1], md_co = F.chunk_off[0];
    for (int k = 0; k < F.nblocks; ++k) {
        const unsigned long long t_blk = (timers_on && blockIdx.x == 0) ? jw_globaltimer() : 0ull;
        const int64_t s = md_s;
        const int b = (int)(md_e - md_s);
        const int64_t chunk_off_k = md_co;
        md_s = md_e;
        if (k + 1 < F.nblocks) { md_e = F.C.starts[k + 2]; md_co = F.chunk_off[k + 1]; }
        if (is_stream_cta) {
        int prev_cnt = 0;
        const int ap = k - 1 - lag;              // block whose updates reach ycorr before this block streams
        if (ap >= 0 && !pipe) {
            if (tid == 0) s_ok = jw_spin_ge(F.done, ap + 1, F.flags) ? 1 : 0;
            __syncthreads();
            if (!s_ok) return;
            prev_cnt = __ldcg(&F.act_cnt_blk[ap]);
        }
        JW_PHASE(0);
        const int nchunks = (b + 15) >> 4;
        const int32_t* prev_idx = F.act_idx_all + (ap >= 0 ? F.C.starts[ap] : 0);
        bool rebuild = (k == 0) || prev_cnt > 0 || !single;
        long long sq_blk[T];
#pragma unroll
        for (int kk = 0; kk < T; ++kk) sq_blk[kk] = 0;

        for (int vs = F.vs0 + blockIdx.x; vs < F.vs1; vs += n_stream) {
            const int64_t row0 = (int64_t)vs * R;
            float v[T];
#pragma unroll
            for (int kk = 0; kk < T; ++kk) v[kk] = 0.0f;
            const bool from_records = (MODE == 1) && ap >= 0;
            if (gather_mode) {
                prev_cnt = s_gcnt[k & 1];                // written before the barrier that ended the previous block
                rebuild = (k == 0) || prev_cnt > 0;
            }
            if constexpr (MODE == 1) { if (from_records) {
                // ---- (1a) pipelined chain: block ap's commits arrive as records; every row of the slice
                //      replays them in commit order against its own genotypes.  The bytes come from this CTA's
                //      own tile of block ap (streamed two panels ago, still in L2). ----
                bool okr = true;
                int cnt = 0;
                if (tid < R) {
                    const int64_t row = row0 + tid;
                    const bool rv = row < nloc;
#pragma unroll
                    for (int kk = 0; kk < T; ++kk) v[kk] = rv ? ycorr_l[kk * n + row] : 0.0f;
                    const int sh = (tid & 3) << 1;
#ifdef JW_NEXT_META
                    // uniform panels (every block uniform_b markers, the last one possibly shorter): the block's
                    // metadata is arithmetic, so the record words are requested at once instead of one dependent
                    // L2 round trip later (profiles/r1_fused_kernel_stall_hotspots.md, item 1)
                    const bool uni = F.uniform_b > 0;
                    const int64_t s_ap = uni ? (int64_t)ap * F.uniform_b : F.C.starts[ap];
                    const int64_t e_ap = uni ? min(p, s_ap + F.uniform_b) : F.C.starts[ap + 1];
                    const int nch_ap = ((int)(e_ap - s_ap) + 15) >> 4;
                    const int64_t co_ap = uni ? (int64_t)ap * ((F.uniform_b + 15) >> 4) : F.chunk_off[ap];
                    const int upb = (F.uniform_b + JW_CHAIN_SB - 1) / JW_CHAIN_SB;
                    const int u0_ap = uni ? ap * upb : F.P.blk_unit0[ap];
                    const int u1_ap = uni ? u0_ap + ((int)(e_ap - s_ap) + JW_CHAIN_SB - 1) / JW_CHAIN_SB : F.P.blk_unit0[ap + 1];
#else
                    const int64_t s_ap = F.C.starts[ap];
                    const int nch_ap = ((int)(F.C.starts[ap + 1] - s_ap) + 15) >> 4;
                    const int64_t co_ap = F.chunk_off[ap];
                    const int u0_ap = F.P.blk_unit0[ap], u1_ap = F.P.blk_unit0[ap + 1];
#endif
                    const uint8_t* tile_ap = F.tiled +
                        ((size_t)(co_ap * F.n_vs + (int64_t)vs * nch_ap) * Gs) * 16 + ((size_t)(tid >> 2) << 4);
                    okr = jw_rec_foreach<T>(F.P, u0_ap, u1_ap,
                                            [&](const int us_, const int nv, const jw_rec_reader<T>& RR) {
#ifdef JW_NEXT_META
                        const int pbase = (us_ - u0_ap) * JW_CHAIN_SB;    // units are cut every JW_CHAIN_SB markers of a panel
#else
                        const int pbase = (int)(F.P.unit_start[us_] - s_ap);
#endif
                        unsigned bytes[JW_REC_BATCH]; float mus[JW_REC_BATCH];
#pragma unroll
                        for (int q = 0; q < JW_REC_BATCH; ++q) {
                            bytes[q] = 0; mus[q] = 0.0f;
                            if (q < nv && rv) {
                                const int pm = pbase + RR.code(q);
                                bytes[q] = tile_ap[((size_t)(pm >> 4) * Gs << 4) + (pm & 15)];
                                mus[q] = F.C.means[s_ap + pm];
                            }
                        }
#pragma unroll
                        for (int q = 0; q < JW_REC_BATCH; ++q) {
                            if (q < nv && rv) {
                                const unsigned code = (bytes[q] >> sh) & 3u;
                                const float xv = (code == 3u ? mus[q] : (float)code) - mus[q];
#pragma unroll
                                for (int kk = 0; kk < T; ++kk) {
                                    const float d = RR.d(q, kk);
                                    if (d != 0.0f) v[kk] = fmaf(d, xv, v[kk]);
                                }
                            }
                        }
                        cnt += nv;
                    });
                    if (tid == 0) s_ok = cnt;
                }
                if (__syncthreads_or(okr ? 0 : 1)) return;
                prev_cnt = s_ok;
                rebuild = (k == 0) || prev_cnt > 0 || !single;
                JW_PHASE(0);
            } }
            if (rebuild) {
                // ---- (1) fused axpy of the previous block + fixed-point image of the slice ----
                long long qs[T];
#pragma unroll
                for (int kk = 0; kk < T; ++kk) qs[kk] = 0;
                if (tid < R) {
                    const int64_t row = row0 + tid;
                    const bool rv = row < nloc;
                    if (gather_mode) {
#pragma unroll
                        for (int kk = 0; kk < T; ++kk) {
                            v[kk] = (k == 0) ? (rv ? ycorr_l[kk * n + row] : 0.0f) : s_yn[kk * R + tid];
                            s_y[kk * R + tid] = v[kk];
                        }
                    } else if (!from_records) {
#pragma unroll
                        for (int kk = 0; kk < T; ++kk) v[kk] = rv ? ycorr_l[kk * n + row] : 0.0f;
                    }
                    if (rv && prev_cnt > 0) {
                        if (!from_records && !gather_mode) {
                            const int sh = (int)(row & 3) << 1;
                            const int64_t byte = row >> 2;
                            for (int a = 0; a < prev_cnt; ++a) {
                                const int64_t j = __ldcg(prev_idx + a);
                                const unsigned code = (F.packed[j * F.stride_d + byte] >> sh) & 3u;
                                const float mu = F.C.means[j];
                                const float xv = (code == 3u ? mu : (float)code) - mu;
#pragma unroll
                                for (int kk = 0; kk < T; ++kk) {
                                    const float d = __ldcg(&F.C.dalpha[kk * p + j]);
                                    if (d != 0.0f) v[kk] = fmaf(d, xv, v[kk]);
                                }
                            }
                        }
#pragma unroll
                        for (int kk = 0; kk < T; ++kk) ycorr_l[kk * n + row] = v[kk];
                    }
                    int ovf = 0;
#pragma unroll
                    for (int kk = 0; kk < T; ++kk) {
                        const int q = rv ? jw_quantize(v[kk], F.scale, &ovf) : 0;
                        yqs[kk * R + tid] = q;
                        qs[kk] = q;
                    }
                    if (ovf) atomicOr(&F.flags[0], 1);
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
                        if (single) sq_keep[kk] = v; else sq_blk[kk] += v;
                    }
                }
                // ---- (2) lookup tables: entry e of group g = sum over its 4 individuals ----
                for (int item = tid; item < Gs * 16; item += JW_FUSED_THREADS) {
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
            }
            JW_PHASE(1);
            // ---- (3) stream the block's genotypes: one lookup per byte (4 individuals) ----
            const uint8_t* tile = F.tiled +
                ((size_t)(chunk_off_k * F.n_vs + (int64_t)vs * nchunks) * Gs) * 16;
            for (int mc = warp; mc < nchunks && warp < nws; mc += nws) {
                int acc[16][W];
#pragma unroll
                for (int q = 0; q < 16; ++q)
#pragma unroll
                    for (int comp = 0; comp < W; ++comp) acc[q][comp] = 0;
                // all of this chunk's 128-bit loads are issued before the first lookup
                uint4 dv[JW_FUSED_MAX_GS / 32];
#pragma unroll
                for (int gb = 0; gb < JW_FUSED_MAX_GS / 32; ++gb) {
                    const int g = gb * 32 + lane;
                    dv[gb] = make_uint4(0, 0, 0, 0);
                    if (g < Gs) dv[gb] = __ldg(reinterpret_cast<const uint4*>(tile + ((size_t)(mc * Gs + g) << 4)));
                }
#ifdef JW_NEXT_L1PF
                // the warp's next chunk -> L1 (no registers involved), so that its LDG.128 do not expose the L2
                // round trip once per chunk (hotspots, item 2); one 128-byte line per lane
                if (mc + nws < nchunks) {
                    const uint8_t* nx = tile + ((size_t)(mc + nws) * Gs << 4);
                    const uint8_t* line = reinterpret_cast<const uint8_t*>(((unsigned long long)nx & ~127ull)) + ((size_t)lane << 7);
                    if (line < nx + ((size_t)Gs << 4)) asm volatile("prefetch.global.L1 [%0];" :: "l"(line));
                }
#endif
#pragma unroll
                for (int gb = 0; gb < JW_FUSED_MAX_GS / 32; ++gb) {
                    const int g = gb * 32 + lane;
                    if (g < Gs) {
                        const unsigned char* tg = tab + jw_tab_sub<W>(gb);
                        const uint32_t laneoff = (uint32_t)lane * 4u * W;
                        const uint32_t wds[4] = {dv[gb].x, dv[gb].y, dv[gb].z, dv[gb].w};
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            // byte1 = packed byte q, byte0 = lane offset: (byte << 8) | laneoff
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
                    // round 1 on the 32-bit partials (half the shuffles, half the live registers)
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
                    const int jj = mc * 16 + q;
                    if ((lane & 1) == 0 && jj < b && tot != 0) {
                        long long* dst = (MISS && comp == 1) ? &F.mq[s + jj]
                                                             : &F.dq[(int64_t)(MISS ? 0 : comp) * p + s + jj];
#ifdef JW_NEXT_RED
                        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(dst), "l"((unsigned long long)tot) : "memory");
#else
                        atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)tot);
#endif
                    }
                }
            }
            if (timed && blockIdx.x == 0) F.C.counters[38] += jw_globaltimer() - t_blk;   // this warp's chunks are done
            if (timers_on && blockIdx.x == 0 && lane == 0 && warp < nws) atomicMax(&s_tmax, (unsigned)(jw_globaltimer() - t_blk));
            if constexpr (MODE == 2) { if (gather_mode && warp == nws) {
                // ---- gather warp: replay block k-1's commits (in commit order) on this CTA's rows ----
                bool okg = true;
                int cnt = 0;
                const int gp = k - lag;                                // the block whose updates block k+1 starts from
                constexpr int NG = JW_FUSED_MAX_GS / 32;              // consecutive byte groups (4 rows each) per lane
                float gv[NG * 4][T];
#pragma unroll
                for (int i = 0; i < NG * 4; ++i)
#pragma unroll
                    for (int kk = 0; kk < T; ++kk)
                        gv[i][kk] = (NG * 4 * lane + i < R) ? s_y[kk * R + NG * 4 * lane + i] : 0.0f;
                if (gp >= 0) {
                    const int64_t s_gp = F.C.starts[gp];
                    const int nch_gp = ((int)(F.C.starts[gp + 1] - s_gp) + 15) >> 4;
                    // this CTA's tile of block gp (streamed two panels ago); lane's groups are NG*lane .. NG*lane+NG-1
                    const uint8_t* tile_gp = F.tiled +
                        ((size_t)(F.chunk_off[gp] * F.n_vs + (int64_t)vs * nch_gp) * Gs) * 16 + ((size_t)(NG * lane) << 4);
                    okg = jw_rec_foreach<T, true>(F.P, F.P.blk_unit0[gp], F.P.blk_unit0[gp + 1],
                                            [&](const int us_, const int nv, const jw_rec_reader<T>& RR) {
                        const int pbase = (int)(F.P.unit_start[us_] - s_gp);
                        // every byte of the batch is requested before the first one is used
                        unsigned bytes[JW_REC_BATCH][NG]; float mus[JW_REC_BATCH];
#pragma unroll
                        for (int q = 0; q < JW_REC_BATCH; ++q) {
                            mus[q] = 0.0f;
#pragma unroll
                            for (int jg = 0; jg < NG; ++jg) bytes[q][jg] = 0u;
                            if (q < nv) {
                                const int pm = pbase + RR.code(q);
                                mus[q] = F.C.means[s_gp + pm];
                                const uint8_t* col = tile_gp + ((size_t)(pm >> 4) * Gs << 4) + (pm & 15);
#pragma unroll
                                for (int jg = 0; jg < NG; ++jg) if (NG * lane + jg < Gs) bytes[q][jg] = col[(size_t)jg << 4];
                            }
                        }
#pragma unroll
                        for (int q = 0; q < JW_REC_BATCH; ++q) {
                            if (q < nv) {
#pragma unroll
                                for (int i = 0; i < NG * 4; ++i) {
                                    const unsigned code = (bytes[q][i >> 2] >> ((i & 3) << 1)) & 3u;
                                    const float xv = (code == 3u ? mus[q] : (float)code) - mus[q];
                                    const bool rowv = (NG * 4 * lane + i < R) && (row0 + NG * 4 * lane + i < nloc);
#pragma unroll
                                    for (int kk = 0; kk < T; ++kk) {
                                        const float d = RR.d(q, kk);
                                        if (d != 0.0f && rowv) gv[i][kk] = fmaf(d, xv, gv[i][kk]);
                                    }
                                }
                            }
                        }
                        cnt += nv;
                    });
                }
#pragma unroll
                for (int i = 0; i < NG * 4; ++i)
#pragma unroll
                    for (int kk = 0; kk < T; ++kk) if (NG * 4 * lane + i < R) s_yn[kk * R + NG * 4 * lane + i] = gv[i][kk];
                if (lane == 0) s_gcnt[(k + 1) & 1] = cnt;  // read by every thread at the start of block k+1
                if (timers_on && lane == 0 && blockIdx.x == 0 && F.C.counters != nullptr) F.C.counters[37] += jw_globaltimer() - t_blk;
                if (!okg) gather_failed = true;
            } }
            if (!single) __syncthreads();       // the next slice overwrites the tables
        }
        // ---- (4) publish this CTA's contribution, then CTA 0 runs the chain ----
        if (__syncthreads_or(gather_failed ? 1 : 0)) return;
        if (timed && blockIdx.x == 0) {
            F.C.counters[39] += jw_globaltimer() - t_blk;        // end-of-block barrier passed
            F.C.counters[47] += s_tmax; s_tmax = 0;               // last streaming warp done
        }
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < T; ++kk) {
                const long long v = single ? sq_keep[kk] : sq_blk[kk];
                if (v != 0) atomicAdd(reinterpret_cast<unsigned long long*>(&F.sq_acc[k * T + kk]), (unsigned long long)v);
            }
            // release at gpu scope: cumulative over the CTA's atomics (ordered before by the barrier)
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" :: "l"(&F.arrive[k]) : "memory");
            if (timed && blockIdx.x == 0) F.C.counters[45] += jw_globaltimer() - t_blk;   // arrive published
        }
        if (warp == 1 && k + 1 < F.nblocks && F.l2_prefetch) {
            // while the chain runs: pull the next block's tile(s) of this CTA into L2
            const int nb1 = (int)(F.C.starts[k + 2] - F.C.starts[k + 1]);
            const int nch1 = (nb1 + 15) >> 4;
            for (int vs = F.vs0 + blockIdx.x; vs < F.vs1; vs += n_stream) {
                const uint8_t* t1 = F.tiled + ((size_t)(F.chunk_off[k + 1] * F.n_vs + (int64_t)vs * nch1) * Gs) * 16;
                const unsigned total = (unsigned)nch1 * Gs * 16;
                const unsigned per = ((total / 32) + 15) & ~15u;
                const unsigned off = per * lane;
                if (off < total) jw_prefetch_l2(t1 + off, min(per, total - off));
            }
        }
        JW_PHASE(2);
        }   // streaming role
        if constexpr (MULTI) { if (is_comm_cta) {
            // wait for this GPU's slices, then push the block's partial rhs to every rank (own included)
            if (tid == 0) s_ok = jw_spin_ge(&F.arrive[k], n_stream * F.arrive_mult, F.flags) ? 1 : 0;
            __syncthreads();
            if (!s_ok) return;
            jw_comm_push<T>(F.dq, F.C.mq ? F.mq : nullptr, F.sq_acc + k * T, F.peer_slots, F.world,
                            (int64_t)(k & (JW_X_RING - 1)) * F.ring_stride + (int64_t)F.rank * F.slot_stride, F.slot_b,
                            F.tag_base + (unsigned)k + 1u, p, s, b);
            const bool okp = jw_comm_pull<T>(F.dq, F.C.mq ? F.mq : nullptr, F.sq_acc + k * T, F.my_slots, F.world,
                                             (int64_t)(k & (JW_X_RING - 1)) * F.ring_stride, F.slot_stride, F.slot_b,
                                             F.tag_base + (unsigned)k + 1u, p, s, b, F.flags);
            if (__syncthreads_or(okp ? 0 : 1)) return;
            // the sums are in place: one more arrival releases the chain (release is cumulative over the CTA's stores)
            if (tid == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" :: "l"(&F.arrive[k]) : "memory");
        } }
        if constexpr (MODE == 0) { if (is_chain_cta) {
            jw_chain_blk B;
            B.xgram = nullptr; B.xgram2 = nullptr; B.xlist = nullptr; B.xcount = nullptr; B.xstart = 0; B.xgram_next = nullptr; B.b_next = 0;
            if (lag && k > 0) {
                B.xgram = F.gramx + F.gramx_off[k];
                B.xlist = F.act_idx_all + F.C.starts[k - 1];
                B.xcount = F.act_cnt_blk + (k - 1);
                B.xstart = F.C.starts[k - 1];
            }
            if (lag && k + 1 < F.nblocks) {
                B.xgram_next = F.gramx + F.gramx_off[k + 1];
                B.b_next = (int)(F.C.starts[k + 2] - F.C.starts[k + 1]);
            }
            B.s = s; B.b = b; B.gram_off = F.C.gram_off[k];
            B.xcount_smem = (lag && F.two_lists) ? prev_commits : -1;
            B.prefetch_s = 0; B.prefetch_b = 0;
            if (k + 1 < F.nblocks) { B.prefetch_s = F.C.starts[k + 1]; B.prefetch_b = (int)(F.C.starts[k + 2] - F.C.starts[k + 1]); }
            B.xslots = nullptr; B.xworld = 1; B.slot_stride = 0; B.slot_b = 0; B.xtag = 0; B.xflags = F.flags;
            B.sq = F.sq_acc + k * T;
            B.act_idx = F.act_idx_all + s;
            B.act_cnt = F.act_cnt_blk + k;
            B.write_active_list = 1;
            auto wait_all = [&]() -> bool {
                if (tid == 0) s_ok = jw_spin_ge(&F.arrive[k], n_stream + (multi ? 1 : 0), F.flags) ? 1 : 0;
                __syncthreads();
                JW_PHASE(3);
                return s_ok != 0;
            };
            unsigned char* chain_smem = reinterpret_cast<unsigned char*>(yqs + 3 * TRp);
            const int nc = jw_chain_block<METHOD, T, false>(F.C, B, k, wait_all, chain_smem, F.list_cap);
            if (nc < 0) return;
            prev_commits = nc;
            __syncthreads();
            if (tid == 0) { __threadfence(); jw_st_release(F.done, k + 1); }
            JW_PHASE(4);
        } }
    }
    if (timed) {
        const int base = (blockIdx.x == 0 && !(is_chain_cta && lag)) ? 32 : (is_chain_cta ? 48 : 40);
        for (int i = 0; i < 5; ++i) F.C.counters[base + i] = ph[i];
    }
#undef JW_PHASE
}

// final axpy of the last block (the in-kernel apply always lags one block behind)
template <int T>
__global__ void __launch_bounds__(256)
jw_k_apply_last(const uint8_t* __restrict__ packed, int64_t stride_d, int64_t n, int64_t p,
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static void jw_fused_free(jwas_handle* h) {
    jw_fused_state* f = (jw_fused_state*)h->fused;
    if (!f) return;
    void* ptrs[] = {f->d_tiled, f->d_chunk_off, f->d_chunk_block, f->d_arrive, f->d_done, f->d_sq_acc, f->d_act_cnt_blk,
                    f->d_unit_start, f->d_unit_blk, f->d_blk_unit0, f->d_rec, f->d_act_cnt_unit};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete f;
    h->fused = nullptr;
}

static bool jw_fused_supported(const jwas_handle* h) {
    if (h->t == 1) return true;
    if (h->t == 2 && !h->has_missing) return true;
    return false;
}

static int jw_fused_prepare(jwas_handle* h) {
    jw_fused_free(h);
    if (!jw_fused_supported(h)) return 0;
    jw_fused_state* f = new jw_fused_state();
    h->fused = f;
    const int64_t nbytes = (jw_nloc(h) + 3) / 4;          // the rows stored on this rank
    f->W = (h->t == 2 || h->has_missing) ? 2 : 1;
    // SMs kept for the chain (lag = 1): one, or chain_ctas of them for the pipelined chain; one more for the
    // NVLink push when rows are sharded
    f->n_chain = h->opt_chain_ctas > 0 ? (int)std::min<int64_t>(h->opt_chain_ctas, std::max(1, h->sm_count / 4)) : 0;
    const int chain_sms = std::max(1, f->n_chain);
    const int64_t streamers = std::max(1, h->sm_count - chain_sms - (h->world > 1 ? 1 : 0));
    int64_t gs = (nbytes + streamers - 1) / streamers;
    if (gs > JW_FUSED_MAX_GS) gs = JW_FUSED_MAX_GS;
    if (gs < 1) gs = 1;
    f->Gs = (int)gs;
    f->TS = (int)((gs + 31) / 32 * 32);
    f->n_vs = (int)((nbytes + gs - 1) / gs);
    f->n_cta = std::min<int>(h->sm_count, f->n_vs);
    // commit lists: one for panels above 1024 markers; two (ping-pong, capacity = largest block) when they
    // fit, so that the lagged schedule's cross-Gram correction reads the previous block's commits on chip
    // tables | yq image | this CTA's rows of ycorr, twice (current / after the next replay)
    const size_t base_smem = (size_t)(f->W == 1 ? 2 : 3) * 65536 + 3 * (((size_t)h->t * f->Gs * 4 + 3) & ~(size_t)3) * 4;
    f->two_lists = (base_smem + jw_chain_smem_bytes(h->t, (int)h->maxb, 2) <= 227 * 1024) ? 1 : 0;
    f->list_cap = f->two_lists ? (int)h->maxb : (h->maxb > JW_MAX_BLOCK ? (int)h->maxb : 0);
    f->smem = base_smem + jw_chain_smem_bytes(h->t, f->list_cap, f->two_lists ? 2 : 1);
    // pipelined chain CTAs never stream: their scratch (no commit lists) overlays the tables
    f->smem_pipe = std::max(base_smem, jw_chain_unit_smem_bytes(h->t));
    f->legacy_ok = f->smem <= 227 * 1024;
    const bool pipe_ok = f->n_chain > 0 && f->smem_pipe <= 227 * 1024 - 2048;
    if (!f->legacy_ok && !pipe_ok) { delete f; h->fused = nullptr; return 0; }   // engine 0 only for this shape
    std::vector<int64_t> coff(h->nblocks + 1, 0);
    std::vector<int32_t> cblk;
    for (int64_t k = 0; k < h->nblocks; ++k) {
        int64_t nc = (h->starts[k + 1] - h->starts[k] + 15) / 16;
        coff[k + 1] = coff[k] + nc;
        for (int64_t c = 0; c < nc; ++c) cblk.push_back((int32_t)k);
    }
    f->total_chunks = coff[h->nblocks];
    size_t tiled_bytes = (size_t)f->total_chunks * f->n_vs * f->Gs * 16;
    JW_CUDA(cudaMalloc((void**)&f->d_tiled, tiled_bytes));
    JW_CUDA(cudaMalloc((void**)&f->d_chunk_off, coff.size() * sizeof(int64_t)));
    JW_CUDA(cudaMalloc((void**)&f->d_chunk_block, cblk.size() * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&f->d_arrive, h->nblocks * sizeof(int)));
    JW_CUDA(cudaMalloc((void**)&f->d_done, sizeof(int)));
    JW_CUDA(cudaMalloc((void**)&f->d_sq_acc, (size_t)h->nblocks * h->t * sizeof(long long)));
    JW_CUDA(cudaMalloc((void**)&f->d_act_cnt_blk, h->nblocks * sizeof(int32_t)));
    JW_CUDA(cudaMemcpyAsync(f->d_chunk_off, coff.data(), coff.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaMemcpyAsync(f->d_chunk_block, cblk.data(), cblk.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int64_t units = f->total_chunks * f->n_vs * f->Gs;
    jw_k_tile<<<(unsigned)((units + 255) / 256), 256, 0, h->stream>>>(h->d_packed, h->stride_d, nbytes, h->d_starts,
        f->d_chunk_off, f->d_chunk_block, f->total_chunks, f->Gs, f->n_vs, f->d_tiled);
    h->launches += 1;
    JW_CUDA(cudaGetLastError());
    JW_CUDA(cudaStreamSynchronize(h->stream));
    if (f->n_chain > 0) {
        // chain units: every block cut into pieces of <= JW_CHAIN_SB markers
        std::vector<int32_t> ublk, bu0(h->nblocks + 1, 0);
        f->unit_start.clear();
        for (int64_t k = 0; k < h->nblocks; ++k) {
            bu0[k] = (int32_t)ublk.size();
            for (int64_t m = h->starts[k]; m < h->starts[k + 1]; m += JW_CHAIN_SB) { f->unit_start.push_back(m); ublk.push_back((int32_t)k); }
        }
        bu0[h->nblocks] = (int32_t)ublk.size();
        f->nunits = (int)ublk.size();
        f->unit_start.push_back(h->p);
        f->rec_bytes = (size_t)f->nunits * JW_REC_STRIDE * h->t * sizeof(unsigned long long);
        JW_CUDA(cudaMalloc((void**)&f->d_unit_start, f->unit_start.size() * sizeof(int64_t)));
        JW_CUDA(cudaMalloc((void**)&f->d_unit_blk, ublk.size() * sizeof(int32_t)));
        JW_CUDA(cudaMalloc((void**)&f->d_blk_unit0, bu0.size() * sizeof(int32_t)));
        JW_CUDA(cudaMalloc((void**)&f->d_rec, f->rec_bytes));
        JW_CUDA(cudaMalloc((void**)&f->d_act_cnt_unit, f->nunits * sizeof(int32_t)));
        JW_CUDA(cudaMemcpyAsync(f->d_unit_start, f->unit_start.data(), f->unit_start.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
        JW_CUDA(cudaMemcpyAsync(f->d_unit_blk, ublk.data(), ublk.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        JW_CUDA(cudaMemcpyAsync(f->d_blk_unit0, bu0.data(), bu0.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        JW_CUDA(cudaMemsetAsync(f->d_rec, 0, f->rec_bytes, h->stream));          // tag 0 = never written
        JW_CUDA(cudaStreamSynchronize(h->stream));
        f->rec_tag = 0;
    }
    f->ready = true;
    return 0;
}

template <int METHOD, int T, int W, int MODE>
static int jw_fused_launch_mode(jwas_handle* h, jw_fused_state* f, jw_fused_args& F) {
    auto kern = F.world > 1 ? jw_k_fused<METHOD, T, W, MODE, true> : jw_k_fused<METHOD, T, W, MODE, false>;
    const size_t smem = MODE == 3 ? std::max<size_t>(jw_chain_unit_smem_bytes(T), (size_t)JW_WS_SMEM)
                                  : (F.P.n_chain > 0 ? f->smem_pipe : f->smem);
    JW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    JW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, JW_FUSED_THREADS, smem));
    JW_REQUIRE(occ >= 1, "fused sweep kernel does not fit on an SM");
    void* args[] = {(void*)&F};
    const int chain_sms = std::max(1, F.P.n_chain);
    const int grid = F.world > 1 ? std::min<int>(h->sm_count, (F.vs1 - F.vs0) + 1 + chain_sms)
                                 : (F.lag ? std::min<int>(h->sm_count, f->n_vs + chain_sms) : f->n_cta);
    JW_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(JW_FUSED_THREADS), args, smem, h->stream));
    h->launches += 1;
    return 0;
}
template <int METHOD, int T, int W>
static int jw_fused_launch(jwas_handle* h, jw_fused_state* f, jw_fused_args& F) {
    if (F.P.n_chain <= 0) {
        JW_REQUIRE(f->legacy_ok, "engine 1: this panel size needs the pipelined chain (option chain_ctas >= 1 with lag = 1)");
        return jw_fused_launch_mode<METHOD, T, W, 0>(h, f, F);
    }
    // the gather warp and the warp-specialised role need one slice per streaming CTA
    const bool single = (F.vs1 - F.vs0) <= h->sm_count - F.P.n_chain - (F.world > 1 ? 1 : 0);
    if constexpr (T == 1 && W == 1) {
        if (h->opt_ws && single) { F.arrive_mult = JW_WS_NS; return jw_fused_launch_mode<METHOD, T, W, 3>(h, f, F); }
    }
    if (F.gather && single) return jw_fused_launch_mode<METHOD, T, W, 2>(h, f, F);
    return jw_fused_launch_mode<METHOD, T, W, 1>(h, f, F);
}

static int jw_fused_sweep(jwas_handle* h, const jw_chain_args& A, float scale) {
    jw_fused_state* f = (jw_fused_state*)h->fused;
    JW_REQUIRE(f && f->ready, "engine 1 (persistent fused sweep) does not support this trait/missing combination");
    const int t = h->t;
    JW_CUDA(cudaMemsetAsync(h->d_dq, 0, (size_t)t * h->p * sizeof(long long), h->stream));
    if (h->has_missing) JW_CUDA(cudaMemsetAsync(h->d_mq, 0, (size_t)t * h->p * sizeof(long long), h->stream));
    JW_CUDA(cudaMemsetAsync(f->d_arrive, 0, h->nblocks * sizeof(int), h->stream));
    JW_CUDA(cudaMemsetAsync(f->d_done, 0, sizeof(int), h->stream));
    JW_CUDA(cudaMemsetAsync(f->d_sq_acc, 0, (size_t)h->nblocks * t * sizeof(long long), h->stream));
    JW_CUDA(cudaMemsetAsync(f->d_act_cnt_blk, 0, h->nblocks * sizeof(int32_t), h->stream));
    JW_CUDA(cudaMemsetAsync(h->d_flags + 2, 0, sizeof(int32_t), h->stream));
    jw_fused_args F;
    F.C = A;
    F.tiled = f->d_tiled; F.chunk_off = f->d_chunk_off;
    F.packed = h->d_packed; F.stride_d = h->stride_d;
    F.Gs = f->Gs; F.TS = f->TS; F.n_vs = f->n_vs; F.nblocks = (int)h->nblocks; F.list_cap = f->list_cap;
    F.lag = A.nreps_mode ? 0 : (int)h->opt_lag; F.timers = (int)h->opt_timers; F.two_lists = f->two_lists;
    JW_REQUIRE(h->gramx_built >= F.lag, "the lagged schedule needs the cross-Gram blocks (set the lag option before jwas_set_blocks or again after it)");
    JW_REQUIRE(F.lag < 2 || f->n_chain > 0, "lag = 2 needs the pipelined chain (option chain_ctas >= 1)");
    F.gramx = h->d_gramx[0]; F.gramx_off = h->d_gramx_off[0];
    F.gramx2 = F.lag >= 2 ? h->d_gramx[1] : nullptr; F.gramx2_off = F.lag >= 2 ? h->d_gramx_off[1] : nullptr;
    F.world = h->world; F.rank = h->rank; F.vs0 = 0; F.vs1 = f->n_vs;     // every slice of the tile is this rank's
    F.row_off = h->row_begin; F.nloc = jw_nloc(h);
    F.peer_slots = nullptr; F.my_slots = nullptr;
    F.slot_stride = 0; F.slot_b = 0; F.ring_stride = 0; F.tag_base = 0;
    if (h->world > 1) {
        JW_REQUIRE(h->ipc_ready && F.lag, "multi-GPU fused sweep needs lag >= 1 and jwas_ipc_import");
        JW_REQUIRE(h->maxb <= h->x_slot_b, "exchange slots are smaller than the largest block (call jwas_ipc_export after jwas_set_blocks)");
        F.peer_slots = (uint4* const*)h->d_peer_slots;
        F.my_slots = (const uint4*)h->d_xbuf;
        F.slot_b = h->x_slot_b; F.slot_stride = h->x_slot_words; F.ring_stride = 8 * h->x_slot_words;   // 8 source ranks per ring entry
        // one tag per (sweep, block), never 0 (the cleared buffer) and never reused within 2^32 blocks
        if (h->sweep_seq * (h->nblocks + 1) + h->nblocks + 2 >= ((int64_t)1 << 32)) {
            JW_CUDA(cudaMemsetAsync(h->d_xbuf, 0, h->xbuf_bytes, h->stream));
            h->sweep_seq = 0;
        }
        F.tag_base = (unsigned)(h->sweep_seq * (h->nblocks + 1));
        h->sweep_seq += 1;
    }
    F.ycorr = h->d_ycorr; F.scale = scale;
    F.arrive = f->d_arrive; F.done = f->d_done; F.sq_acc = f->d_sq_acc;
    F.act_cnt_blk = f->d_act_cnt_blk; F.act_idx_all = h->d_act_idx;
    F.flags = h->d_flags; F.dq = h->d_dq; F.mq = h->d_mq;
    memset(&F.P, 0, sizeof(F.P));
    // the pipelined chain pays when every streaming CTA owns ONE row slice (n up to ~55,000 rows per GPU): with
    // several slices per CTA the stream dominates, the chain is idle anyway, and replaying the records once
    // per slice costs more than the one-chain-CTA hand-off (400,000 x 61,440: 260 vs 284 sweeps/s)
    const int my_slices = f->n_vs;
    const bool one_slice = my_slices <= h->sm_count - std::max(1, f->n_chain) - (h->world > 1 ? 1 : 0);
    const bool pipe = F.lag && f->n_chain > 0 && (one_slice || !f->legacy_ok || F.lag >= 2);
    F.gather = (int)h->opt_gather; F.l2_prefetch = (int)h->opt_l2_prefetch; F.arrive_mult = 1;
    F.uniform_b = 0;
    if (h->nblocks >= 1) {
        const int64_t b0 = h->starts[1] - h->starts[0];
        bool uni = b0 <= (int64_t)JW_MAX_PANEL;
        for (int64_t k = 1; k < h->nblocks && uni; ++k) {
            const int64_t bk = h->starts[k + 1] - h->starts[k];
            uni = (k + 1 < h->nblocks) ? (bk == b0) : (bk <= b0);
        }
        if (uni) F.uniform_b = (int)b0;
    }
    if (pipe) {
        f->rec_tag += 1;
        if (f->rec_tag > 0xffffu) {              // tags wrapped: forget every old record
            JW_CUDA(cudaMemsetAsync(f->d_rec, 0, f->rec_bytes, h->stream));
            f->rec_tag = 1;
        }
        F.P.n_chain = f->n_chain; F.P.nunits = f->nunits;
        F.P.unit_start = f->d_unit_start; F.P.unit_blk = f->d_unit_blk; F.P.blk_unit0 = f->d_blk_unit0;
        F.P.rec = f->d_rec; F.P.tag = f->rec_tag; F.P.act_cnt_unit = f->d_act_cnt_unit; F.P.flags = h->d_flags;
        F.P.sleep_stream = (unsigned)h->opt_poll_ns_stream; F.P.sleep_chain = (unsigned)h->opt_poll_ns_chain;
        JW_CUDA(cudaMemsetAsync(f->d_act_cnt_unit, 0, f->nunits * sizeof(int32_t), h->stream));
    }
    if (h->opt_profile) {
        cudaEvent_t a, b;
        JW_CUDA(cudaEventCreate(&a)); JW_CUDA(cudaEventCreate(&b));
        h->prof_events.push_back(a); h->prof_events.push_back(b);
        JW_CUDA(cudaEventRecord(a, h->stream));
    }
    int rc = 2;
    const bool ms = h->has_missing != 0;
    if (t == 1 && !ms) {
        if (A.method == 0) rc = jw_fused_launch<0, 1, 1>(h, f, F);
        else if (A.method == 1) rc = jw_fused_launch<1, 1, 1>(h, f, F);
    } else if (t == 1 && ms) {
        if (A.method == 0) rc = jw_fused_launch<0, 1, 2>(h, f, F);
        else if (A.method == 1) rc = jw_fused_launch<1, 1, 2>(h, f, F);
    } else if (t == 2 && !ms && A.method == 2) {
        rc = jw_fused_launch<2, 2, 2>(h, f, F);
    } else if (t == 2 && !ms && A.method == 3) {
        rc = jw_fused_launch<3, 2, 2>(h, f, F);
    } else if (t == 2 && !ms && A.method == 4) {
        rc = jw_fused_launch<4, 2, 2>(h, f, F);
    }
    if (rc == 2) jw_set_error("engine 1: unsupported (method, traits, missing) combination");
    if (rc) return rc;
    if (h->opt_profile) JW_CUDA(cudaEventRecord(h->prof_events.back(), h->stream));
    // the axpy of the last block (and of the one before it under the lagged schedule)
    for (int64_t blk = std::max<int64_t>(0, h->nblocks - 1 - F.lag); blk < h->nblocks; ++blk) {
        // one ordered active list per block, or per chain unit of the block (pipelined chain)
        std::vector<std::pair<int64_t, const int32_t*>> lists;
        if (pipe) {
            for (int u = 0; u < f->nunits; ++u)
                if (f->unit_start[u] >= h->starts[blk] && f->unit_start[u] < h->starts[blk + 1])
                    lists.push_back({f->unit_start[u], f->d_act_cnt_unit + u});
        } else lists.push_back({h->starts[blk], f->d_act_cnt_blk + blk});
        for (auto& L : lists) {
            unsigned g = (unsigned)std::max<int64_t>(1, (h->row_end - h->row_begin + 255) / 256);
            if (t == 1)
                jw_k_apply_last<1><<<g, 256, 0, h->stream>>>(jw_packed_g(h), h->stride_d, h->n, h->p, h->d_means, h->d_dalpha,
                    h->d_act_idx + L.first, L.second, h->d_ycorr, h->row_begin, h->row_end);
            else
                jw_k_apply_last<2><<<g, 256, 0, h->stream>>>(jw_packed_g(h), h->stride_d, h->n, h->p, h->d_means, h->d_dalpha,
                    h->d_act_idx + L.first, L.second, h->d_ycorr, h->row_begin, h->row_end);
            h->launches += 1;
            JW_CUDA(cudaGetLastError());
        }
    }
    // abort flag -> error
    int32_t hf[4];
    JW_CUDA(cudaMemcpyAsync(hf, h->d_flags, sizeof(hf), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    if (hf[2]) { jw_set_error("engine 1: inter-CTA wait timed out (sweep aborted)"); return 12; }
    return 0;
}
