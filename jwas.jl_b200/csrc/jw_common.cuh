// jw_common.cuh -- handle layout, error plumbing and small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/jwas_b200.h"
#include "../../include/jwas_contract.h"

#define JW_MAX_TRAITS 4
#define JW_MAX_BLOCK 1024      // chain threads per CTA (one marker each per sub-block)
#define JW_MAX_PANEL 4096      // largest block of the exact schedule (walked in sub-blocks)
#define JW_MAX_CLASSES 8
#define JW_MAX_LAG 2            // deepest lagged exact schedule on the device (the oracle goes to 3)
#define JW_R_CLASSES 4          // BayesR mixture classes on the device (BAYESR_GAMMA, JWAS.jl:12)

void jw_set_error(const std::string& s);

#define JW_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            jw_set_error(std::string(#call) + ": " + cudaGetErrorString(e__));          \
            return 10;                                                                  \
        }                                                                               \
    } while (0)

#define JW_REQUIRE(cond, msg)                                                           \
    do {                                                                                \
        if (!(cond)) { jw_set_error(msg); return 2; }                                   \
    } while (0)

struct jwas_handle {
    int device = 0;
    int64_t n = 0, p = 0, stride = 0, stride_d = 0;   // n: ALL individuals; stride_d: device column pitch (16B multiple)
                                                      // of the rows this rank stores, [row_begin, row_end)
    int t = 1;
    int has_missing = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_sweep_ms = 0.0;
    int64_t launches = 0;

    // genotypes (marker-major .jgb2 image) and per-marker statistics
    uint8_t* d_packed = nullptr;
    float* d_means = nullptr;
    float* d_xpx = nullptr;
    int32_t* d_colsum = nullptr;   // sum of codes over observed rows
    int32_t* d_nvalid = nullptr;   // observed rows
    int32_t* d_cnt = nullptr;      // 3*p: counts of codes 1, 2, 3 over this rank's rows (summed over ranks before use)
    int stats_ready = 0;           // means / xpx / colsum / nvalid are final (after the all-reduce when rows are sharded)
    int ext_means = 0;             // means were supplied by the host (jwas_set_marker_means)

    // sampler state
    float* d_ycorr = nullptr;      // t * n
    float* d_alpha = nullptr;      // t * p
    float* d_beta = nullptr;       // t * p
    int32_t* d_delta = nullptr;    // t * p
    float* d_mean_alpha = nullptr; // posterior accumulators, t * p each
    float* d_mean_alpha2 = nullptr;
    float* d_mean_delta = nullptr;

    // hyper-parameter vectors resident on the device
    double* d_ve = nullptr;        // p (or p*t*t for per-marker G)
    double* d_pi = nullptr;        // p  (BayesR per-marker: p*nclasses; MT per-marker: p*2^t)
    double* d_prep = nullptr;      // 6*p chain constants (BayesABC, repetition 0)
    float* d_prep_beta0 = nullptr; // p
    double* d_draws = nullptr;     // 2*t*p: draws of repetition 0 (BayesR / multi-trait)
    size_t cap_draws = 0;
    double* d_prep_rm = nullptr;   // BayesR / multi-trait rhs-independent terms
    size_t cap_prep_rm = 0;
    double* d_u = nullptr;         // replay tables (allocated on demand)
    double* d_z = nullptr;
    size_t cap_ve = 0, cap_pi = 0, cap_u = 0, cap_z = 0;

    // block partition and Gram blocks
    std::vector<int64_t> starts;   // nblocks+1, 0-based
    std::vector<int64_t> gram_off; // float offset of each block's b*b Gram
    int64_t* d_starts = nullptr;
    int64_t* d_gram_off = nullptr;
    float* d_gram = nullptr;
    // cross-Gram X_{k-d}'X_k towards the d-th previous block, d = 1..lag (lagged schedule), on demand
    float* d_gramx[2] = {nullptr, nullptr};
    int64_t* d_gramx_off[2] = {nullptr, nullptr};
    std::vector<int64_t> gramx_off[2];
    int gramx_built = 0;
    int64_t nblocks = 0, maxb = 0;

    // sweep workspace
    int32_t* d_yq = nullptr;       // t * n fixed-point image of ycorr
    long long* d_sq = nullptr;     // t
    long long* d_dq = nullptr;     // t * p  (indexed by marker)
    long long* d_mq = nullptr;     // t * p
    float* d_dalpha = nullptr;     // t * p  net (old - new) per marker of the current block(s)
    int32_t* d_act_idx = nullptr;  // p   ordered list of markers with any dalpha != 0
    int32_t* d_act_cnt = nullptr;  // 1
    int32_t* d_flags = nullptr;    // [0] overflow
    unsigned long long* d_counters = nullptr; // [0] n_active, [1] n_rounds
    float* d_maxabs = nullptr;     // 1 (as uint bits)
    double* d_stats = nullptr;     // reduction scratch
    double* d_partials = nullptr;
    size_t cap_partials = 0;

    // options
    int64_t opt_profile = 0;       // 1 = time the genotype-streaming kernel(s) with CUDA events
    std::vector<cudaEvent_t> prof_events;
    double prof_ms = 0.0; int64_t prof_launches = 0;
    int64_t opt_gram_popc = 0;     // 1 = build Gram blocks with the popcount kernel (default: bf16 tensor-core GEMM)
    int64_t opt_timers = 0;        // 1 = in-kernel phase timers (tools/phase_probe.py)
    int64_t opt_lag = 0;           // L = lagged exact schedule (engine 1): the chains of blocks k-L..k-1 overlap the stream of block k
    int64_t opt_engine = 0;        // 0 = multi-kernel engine, 1 = persistent fused kernel
    int64_t opt_gather = 0;        // pipelined chain: 1 = a gather warp per streaming CTA replays the records under the
                                   // stream (pays off with panels that are a multiple of 31*16 markers), 0 = in line
    int64_t opt_chain_ctas = 0;    // engine 1, lag 1: chain CTAs of the pipelined chain (0 = one-CTA chain)
    int64_t opt_stream_variant = 1;// streamed block rhs (independent schedule): 0 = 512 thr + register double buffer,
                                   // 1 = 1024 thr (measured best: 2.27 ms at cfg2), 2 = 768 thr + double buffer, 3 = 1024 thr + double buffer
    int64_t opt_stream_pf = 0;     // L2 prefetch distance of the streamed block rhs, in chunk iterations (0 = off: measured best)
    int64_t opt_poll_ns_stream = 0, opt_poll_ns_chain = 0;     // back-off after an empty record poll (measured: no effect)
    int64_t opt_ws = 1;            // engine 1, pipelined chain, one trait without missing calls: warp-specialised streaming role
                                   // (builder warps + streaming warps, two table sets; jw_fused_ws.cuh)
    int64_t opt_l2_prefetch = 0;   // engine 1: pull the next panel's tile into L2 at the end of a panel (measured: slower, off)
    // row-sharded multi-GPU sweep: this rank STORES and streams rows [row_begin, row_end) of every column
    // (row_begin is a multiple of 64); ycorr and the sampler state are replicated
    int64_t row_begin = 0, row_end = 0;
    float* d_gath = nullptr; size_t cap_gath = 0;     // all-gather staging (world * chunk * t floats, twice)
    int world = 1, rank = 0;
    void* nccl_comm = nullptr;
    std::vector<int64_t> shard_bounds;   // world+1 row boundaries (multiples of 16 except the last)
    // fused multi-GPU exchange (IPC-mapped peer memory): one allocation = [4 rings][8 source ranks][slot of 16-byte words]
    unsigned char* d_xbuf = nullptr; size_t xbuf_bytes = 0; int64_t x_slot_words = 0; int x_slot_b = 0;
    void* peer_bufs[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    long long** d_peer_slots = nullptr;
    int ipc_ready = 0; int64_t sweep_seq = 0;
    void* fused = nullptr;         // jw_fused_state (engine 1)
    float next_maxabs = -1.0f;     // carried from the previous sweep's stats when ycorr untouched
};

// number of rows stored on this rank, and the packed image addressed by GLOBAL row (byte row>>2 of a column)
static inline int64_t jw_nloc(const jwas_handle* h) { return h->row_end - h->row_begin; }
static inline const uint8_t* jw_packed_g(const jwas_handle* h) { return h->d_packed - (h->row_begin >> 2); }

__device__ __forceinline__ unsigned jw_dcode(const uint8_t* col, int64_t i) {
    return (col[i >> 2] >> ((i & 3) << 1)) & 3u;
}
