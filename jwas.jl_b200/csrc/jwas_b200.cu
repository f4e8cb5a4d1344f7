// jwas_b200.cu -- C ABI of libjwasb200.so (see include/jwas_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false ...
//        (--fmad=false is part of the arithmetic contract: no implicit contraction).
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cmath>
#include "jw_common.cuh"
#include "jw_setup_kernels.cuh"
#include "jw_sweep_kernels.cuh"
#include "jw_fused_sweep.cuh"
#include "jw_stream_kernel.cuh"
#include "jw_nccl.cuh"
#include <cublas_v2.h>

static thread_local std::string g_err;
void jw_set_error(const std::string& s) { g_err = s; }
extern "C" const char* jwas_last_error(void) { return g_err.c_str(); }

extern "C" int jwas_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

#define JW_LAUNCH_CHECK(h)                                                              \
    do {                                                                                \
        (h)->launches += 1;                                                             \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            jw_set_error(std::string("kernel launch failed: ") + cudaGetErrorString(e__)); \
            return 11;                                                                  \
        }                                                                               \
    } while (0)

extern "C" int jwas_destroy(jwas_handle* h);
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T>
static int ensure_cap(T** ptr, size_t* cap, size_t need) {
    if (*cap >= need && *ptr) return 0;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    JW_CUDA(cudaMalloc((void**)ptr, need * sizeof(T)));
    *cap = need;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Row shards: words of 64 individuals (16 packed bytes) split evenly, the remainder to the first ranks.
static void shard_range(int64_t n, int rank, int world, int64_t* b, int64_t* e) {
    const int64_t nw = ceil_div(n, 64);
    auto bound = [&](int r) { return std::min<int64_t>(n, (nw / world * r + std::min<int64_t>(r, nw % world)) * 64); };
    *b = bound(rank); *e = rank + 1 == world ? n : bound(rank + 1);
}
extern "C" int jwas_shard_range(int64_t n_obs, int rank, int world, int64_t* begin, int64_t* end) {
    JW_REQUIRE(begin && end, "jwas_shard_range: null argument");
    JW_REQUIRE(n_obs > 0 && world >= 1 && rank >= 0 && rank < world, "jwas_shard_range: bad arguments");
    shard_range(n_obs, rank, world, begin, end);
    return 0;
}

static int create_common(int64_t n, int64_t p, int t, int64_t row_begin, int64_t row_end, int device, jwas_handle** out) {
    JW_REQUIRE(out != nullptr, "jwas_create: out is NULL");
    *out = nullptr;
    JW_REQUIRE(n > 0 && p > 0, "Genotype data is empty.");
    JW_REQUIRE(p < (int64_t)2147483647, "jwas_create: too many markers for 32-bit marker ids");
    JW_REQUIRE(t >= 1 && t <= JW_MAX_TRAITS, "jwas_create: number of traits must be 1..4");
    JW_REQUIRE(row_begin >= 0 && row_begin < row_end && row_end <= n && (row_begin & 63) == 0,
               "jwas_create: the row range must lie inside 0..nObs and begin on a multiple of 64");
    int ndev = jwas_device_count();
    JW_REQUIRE(ndev > 0, "no CUDA device is visible: libjwasb200 has no CPU fallback");
    JW_REQUIRE(device >= 0 && device < ndev, "jwas_create: device index out of range");
    JW_CUDA(cudaSetDevice(device));

    jwas_handle* h = new jwas_handle();
    h->device = device; h->n = n; h->p = p; h->t = t;
    h->row_begin = row_begin; h->row_end = row_end;
    const int64_t nloc = row_end - row_begin;
    h->stride = (nloc + 3) / 4;
    h->stride_d = ceil_div((nloc + 3) / 4, 16) * 16;
    cudaDeviceProp prop;
    JW_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    JW_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    JW_CUDA(cudaEventCreate(&h->ev0));
    JW_CUDA(cudaEventCreate(&h->ev1));

    size_t tp = (size_t)t * p, tn = (size_t)t * n;
    JW_CUDA(cudaMalloc((void**)&h->d_packed, (size_t)p * h->stride_d));
    JW_CUDA(cudaMemsetAsync(h->d_packed, 0, (size_t)p * h->stride_d, h->stream));
    JW_CUDA(cudaMalloc((void**)&h->d_means, p * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_xpx, p * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_colsum, p * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_nvalid, p * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_cnt, (size_t)3 * p * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_ycorr, tn * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_alpha, tp * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_beta, tp * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_delta, tp * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_mean_alpha, tp * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_mean_alpha2, tp * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_mean_delta, tp * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_yq, tn * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_sq, JW_MAX_TRAITS * sizeof(long long)));
    JW_CUDA(cudaMalloc((void**)&h->d_dq, tp * sizeof(long long)));
    JW_CUDA(cudaMalloc((void**)&h->d_mq, tp * sizeof(long long)));
    JW_CUDA(cudaMalloc((void**)&h->d_dalpha, tp * sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_act_idx, p * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_act_cnt, sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_flags, 4 * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_counters, 64 * sizeof(unsigned long long)));
    JW_CUDA(cudaMalloc((void**)&h->d_maxabs, sizeof(float)));
    JW_CUDA(cudaMalloc((void**)&h->d_stats, 64 * sizeof(double)));
    for (void* z : {(void*)h->d_ycorr, (void*)h->d_yq}) JW_CUDA(cudaMemsetAsync(z, 0, tn * 4, h->stream));
    for (void* z : {(void*)h->d_alpha, (void*)h->d_beta, (void*)h->d_delta, (void*)h->d_mean_alpha,
                    (void*)h->d_mean_alpha2, (void*)h->d_mean_delta, (void*)h->d_dalpha})
        JW_CUDA(cudaMemsetAsync(z, 0, tp * 4, h->stream));
    JW_CUDA(cudaMemsetAsync(h->d_flags, 0, 4 * sizeof(int32_t), h->stream));
    *out = h;
    return 0;
}

// per-marker statistics (GibbsMats: xpRinvx, tools4genotypes.jl:28-36, 247-250): integer counts over the
// rows stored here, summed over the ranks when rows are sharded, then one closed form per marker
static int marker_counts(jwas_handle* h) {
    jw_k_marker_counts<<<(unsigned)ceil_div(h->p, 8), 256, 0, h->stream>>>(h->d_packed, h->stride_d, h->p, h->d_cnt);
    JW_LAUNCH_CHECK(h);
    return 0;
}
static int marker_finalize(jwas_handle* h) {
    if (h->world > 1) {
        jw_nccl_api* N = jw_nccl();
        JW_REQUIRE(N && h->nccl_comm, "sharded handle without an initialised NCCL communicator");
        JW_NCCL(N->AllReduce(h->d_cnt, h->d_cnt, (size_t)3 * h->p, JW_NCCL_INT32, JW_NCCL_SUM, h->nccl_comm, h->stream));
    }
    int* d_hm = (int*)&h->d_flags[1];
    JW_CUDA(cudaMemsetAsync(d_hm, 0, sizeof(int), h->stream));
    jw_k_marker_finalize<<<(unsigned)ceil_div(h->p, 256), 256, 0, h->stream>>>(
        h->d_cnt, h->n, h->p, h->ext_means, h->d_means, h->d_xpx, h->d_colsum, h->d_nvalid, d_hm);
    JW_LAUNCH_CHECK(h);
    JW_CUDA(cudaMemcpyAsync(&h->has_missing, d_hm, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    h->stats_ready = 1;
    return 0;
}
// full handles finish here; a shard (rows of a larger problem) waits for jwas_init_sharding
static int finish_create(jwas_handle* h) {
    int rc = marker_counts(h);
    if (rc) return rc;
    if (h->row_begin == 0 && h->row_end == h->n) return marker_finalize(h);
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int jwas_create_shard(int64_t n, int64_t p, int t, int64_t row_begin, int64_t row_end,
                                 const uint8_t* packed_rows, int64_t stride, int device, jwas_handle** out) {
    JW_REQUIRE(n > 0 && p > 0, "Genotype data is empty.");
    JW_REQUIRE(packed_rows != nullptr, "jwas_create: packed is NULL");
    JW_REQUIRE(row_end > row_begin && stride >= (row_end - row_begin + 3) / 4, "jwas_create: stride_bytes is smaller than cld(rows,4)");
    int rc = create_common(n, p, t, row_begin, row_end, device, out);
    if (rc) return rc;
    jwas_handle* h = *out;
    JW_CUDA(cudaMemcpy2DAsync(h->d_packed, h->stride_d, packed_rows, stride, (row_end - row_begin + 3) / 4, p,
                              cudaMemcpyHostToDevice, h->stream));
    rc = finish_create(h);
    if (rc) { jwas_destroy(h); *out = nullptr; }
    return rc;
}
extern "C" int jwas_create(int64_t n, int64_t p, int t, const uint8_t* packed, int64_t stride,
                           int device, jwas_handle** out) {
    JW_REQUIRE(n > 0 && p > 0, "Genotype data is empty.");
    JW_REQUIRE(stride >= (n + 3) / 4, "jwas_create: stride_bytes is smaller than cld(nObs,4)");
    return jwas_create_shard(n, p, t, 0, n, packed, stride, device, out);
}

// synthetic genotypes: one thread per packed byte (4 individuals); bytes [b0, b0 + nbytes_loc) of every column
__global__ void __launch_bounds__(256)
jw_k_synth(uint8_t* __restrict__ packed, int64_t stride_d, int64_t n, int64_t p, int64_t b0, int64_t nbytes_loc,
           uint64_t seed, uint32_t miss_thr) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nbytes_loc * p) return;
    int64_t j = idx / nbytes_loc, lb = idx % nbytes_loc, b = b0 + lb;
    jw_u32x4 rf = jw_philox4x32_10((uint32_t)j, 0u, 0xF00Du, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    double f = 0.05 + 0.45 * jw_u53(rf.v[0], rf.v[1]);
    uint32_t thr = (uint32_t)(f * 65536.0);
    jw_u32x4 r = jw_philox4x32_10((uint32_t)b, (uint32_t)j, 0xC0DEu, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    jw_u32x4 rm;
    if (miss_thr) rm = jw_philox4x32_10((uint32_t)b, (uint32_t)j, 0xC0DFu, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    unsigned byte = 0;
    for (int k = 0; k < 4; ++k) {
        if (4 * b + k >= n) break;
        uint32_t w = r.v[k];
        unsigned code = ((w & 0xffffu) < thr) + ((w >> 16) < thr);
        if (miss_thr && (rm.v[k] >> 16) < miss_thr) code = 3u;
        byte |= code << (2 * k);
    }
    packed[j * stride_d + lb] = (uint8_t)byte;
}

extern "C" int jwas_create_synthetic_shard(int64_t n, int64_t p, int t, uint64_t seed, double missing_rate,
                                           int64_t row_begin, int64_t row_end, int device, jwas_handle** out) {
    JW_REQUIRE(missing_rate >= 0.0 && missing_rate < 0.5, "missing_rate must be in [0,0.5)");
    int rc = create_common(n, p, t, row_begin, row_end, device, out);
    if (rc) return rc;
    jwas_handle* h = *out;
    const int64_t nbytes_loc = (row_end - row_begin + 3) / 4;
    const int64_t total = nbytes_loc * p;
    // every byte depends on (seed, marker, GLOBAL byte) only: a shard holds exactly the rows of the full matrix
    jw_k_synth<<<(unsigned)ceil_div(total, 256), 256, 0, h->stream>>>(h->d_packed, h->stride_d, n, p, row_begin >> 2,
                                                                     nbytes_loc, seed, (uint32_t)(missing_rate * 65536.0));
    JW_LAUNCH_CHECK(h);
    rc = finish_create(h);
    if (rc) { jwas_destroy(h); *out = nullptr; }
    return rc;
}
extern "C" int jwas_create_synthetic(int64_t n, int64_t p, int t, uint64_t seed, double missing_rate,
                                     int device, jwas_handle** out) {
    return jwas_create_synthetic_shard(n, p, t, seed, missing_rate, 0, n, device, out);
}

// the rows stored on this rank (all rows of an unsharded handle): p columns of cld(rows,4) bytes
extern "C" int jwas_get_packed(jwas_handle* h, uint8_t* packed, int64_t stride) {
    JW_REQUIRE(h && packed, "jwas_get_packed: null argument");
    JW_REQUIRE(stride >= (jw_nloc(h) + 3) / 4, "jwas_get_packed: stride too small");
    JW_CUDA(cudaSetDevice(h->device));
    JW_CUDA(cudaMemcpy2DAsync(packed, stride, h->d_packed, h->stride_d, (jw_nloc(h) + 3) / 4, h->p,
                              cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// centring on means computed elsewhere (the reference centres on ALL genotyped individuals in get_genotypes,
// readgenotypes.jl:372-385, and only then aligns rows to the phenotypes, JWAS.jl:381-402); xpx follows.
extern "C" int jwas_set_marker_means(jwas_handle* h, const float* means) {
    JW_REQUIRE(h && means, "jwas_set_marker_means: null argument");
    JW_REQUIRE(h->nblocks == 0, "jwas_set_marker_means must be called before jwas_set_blocks");
    JW_REQUIRE(h->stats_ready, "jwas_set_marker_means: call jwas_init_sharding first on a sharded handle");
    JW_CUDA(cudaSetDevice(h->device));
    JW_CUDA(cudaMemcpyAsync(h->d_means, means, h->p * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    h->ext_means = 1;
    const int world = h->world; h->world = 1;            // the counts are already global
    int rc = marker_finalize(h);
    h->world = world;
    return rc;
}

extern "C" int jwas_get_gram(jwas_handle* h, int64_t ib, float* out) {
    JW_REQUIRE(h && out, "jwas_get_gram: null argument");
    JW_REQUIRE(ib >= 0 && ib < h->nblocks, "jwas_get_gram: block out of range");
    JW_CUDA(cudaSetDevice(h->device));
    int64_t b = h->starts[ib + 1] - h->starts[ib];
    JW_CUDA(cudaMemcpyAsync(out, h->d_gram + h->gram_off[ib], b * b * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int jwas_destroy(jwas_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    void* ptrs[] = {h->d_packed, h->d_means, h->d_xpx, h->d_colsum, h->d_nvalid, h->d_cnt, h->d_gath, h->d_ycorr, h->d_alpha,
                    h->d_beta, h->d_delta, h->d_mean_alpha, h->d_mean_alpha2, h->d_mean_delta, h->d_ve,
                    h->d_pi, h->d_u, h->d_z, h->d_prep, h->d_prep_beta0, h->d_draws, h->d_prep_rm, h->d_gramx[0], h->d_gramx[1], h->d_gramx_off[0], h->d_gramx_off[1], h->d_starts, h->d_gram_off, h->d_gram, h->d_yq, h->d_sq,
                    h->d_dq, h->d_mq, h->d_dalpha, h->d_act_idx, h->d_act_cnt, h->d_flags, h->d_counters,
                    h->d_maxabs, h->d_stats, h->d_partials};
    for (void* q : ptrs) if (q) cudaFree(q);
    for (int r = 0; r < 8; ++r) if (h->peer_bufs[r]) cudaIpcCloseMemHandle(h->peer_bufs[r]);
    for (void* q : {(void*)h->d_xbuf, (void*)h->d_peer_slots}) if (q) cudaFree(q);
    if (h->nccl_comm && jw_nccl()) jw_nccl()->CommDestroy(h->nccl_comm);
    jw_fused_free(h);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int jwas_get_marker_stats(jwas_handle* h, float* means, float* xpx) {
    JW_REQUIRE(h, "null handle");
    JW_CUDA(cudaSetDevice(h->device));
    if (means) JW_CUDA(cudaMemcpyAsync(means, h->d_means, h->p * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (xpx) JW_CUDA(cudaMemcpyAsync(xpx, h->d_xpx, h->p * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// block partition + Gram blocks (JWAS.jl:73-79 validation; tools4genotypes.jl:259-269)
// ------------------------------------------------------------------------------------------
#define JW_CUBLAS(call)                                                                 \
    do {                                                                                \
        cublasStatus_t r__ = (call);                                                    \
        if (r__ != CUBLAS_STATUS_SUCCESS) {                                             \
            jw_set_error(std::string(#call) + ": cuBLAS status " + std::to_string((int)r__)); \
            return 14;                                                                  \
        }                                                                               \
    } while (0)

// (re)allocates the cross-Gram storage for distances 1..dmax: X_{k-d}' X_k, rows = markers of block k-d
static int alloc_gramx(jwas_handle* h, int dmax) {
    const int64_t nb = h->nblocks;
    for (int d = 1; d <= JW_MAX_LAG; ++d) {
        if (h->d_gramx[d - 1]) { cudaFree(h->d_gramx[d - 1]); h->d_gramx[d - 1] = nullptr; }
        if (h->d_gramx_off[d - 1]) { cudaFree(h->d_gramx_off[d - 1]); h->d_gramx_off[d - 1] = nullptr; }
        h->gramx_off[d - 1].clear();
    }
    h->gramx_built = 0;
    for (int d = 1; d <= dmax; ++d) {
        std::vector<int64_t> xoff(nb, 0);
        int64_t total = 0;
        for (int64_t i = d; i < nb; ++i) { xoff[i] = total; total += (h->starts[i - d + 1] - h->starts[i - d]) * (h->starts[i + 1] - h->starts[i]); }
        h->gramx_off[d - 1] = xoff;
        JW_CUDA(cudaMalloc((void**)&h->d_gramx[d - 1], std::max<size_t>(1, (size_t)total) * sizeof(float)));
        JW_CUDA(cudaMalloc((void**)&h->d_gramx_off[d - 1], nb * sizeof(int64_t)));
        JW_CUDA(cudaMemcpyAsync(h->d_gramx_off[d - 1], xoff.data(), nb * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    }
    return 0;
}

// Gram blocks (and the cross-Gram towards the `cross_max` previous blocks) as bf16 tensor-core GEMMs on unpacked
// 0/1/2 codes; integer-exact (see jw_setup_kernels.cuh).  One pass over the blocks, the previous blocks'
// unpacked panels are kept in a ring for the cross products.
static int build_gram_gemm(jwas_handle* h, int cross_max) {
    // rows stored on this rank; the integer pair counts (exact as FP32 below 2^24) are summed over the ranks
    const int64_t nb = h->nblocks, n = jw_nloc(h);
    const int64_t n_pad = ceil_div(n, 16) * 16;
    jw_nccl_api* N = h->world > 1 ? jw_nccl() : nullptr;
    JW_REQUIRE(h->world == 1 || (N && h->nccl_comm), "sharded handle without an initialised NCCL communicator");
    const int64_t maxb = h->maxb;
    const bool ms = h->has_missing != 0;
    const int NR = cross_max + 1;                       // ring of unpacked panels
    cublasHandle_t cb = nullptr;
    JW_CUBLAS(cublasCreate(&cb));
    JW_CUBLAS(cublasSetStream(cb, h->stream));
    __nv_bfloat16 *C[JW_MAX_LAG + 1] = {nullptr}, *V[JW_MAX_LAG + 1] = {nullptr};
    float* cnt[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int q = 0; q < NR; ++q) {
        JW_CUDA(cudaMalloc((void**)&C[q], (size_t)n_pad * maxb * sizeof(__nv_bfloat16)));
        if (ms) JW_CUDA(cudaMalloc((void**)&V[q], (size_t)n_pad * maxb * sizeof(__nv_bfloat16)));
    }
    for (int q = 0; q < (ms ? 4 : 1); ++q) JW_CUDA(cudaMalloc((void**)&cnt[q], (size_t)maxb * maxb * sizeof(float)));
    if (alloc_gramx(h, cross_max)) return 10;
    const float one = 1.0f, zero = 0.0f;
    auto gemm = [&](const __nv_bfloat16* A, int m, const __nv_bfloat16* B, int nn, float* out) -> cublasStatus_t {
        // out (column-major m x nn, ld m) = A^T (m x n) * B (n x nn): out[c + a*m] = sum_i A[i,c] * B[i,a]
        return cublasGemmEx(cb, CUBLAS_OP_T, CUBLAS_OP_N, m, nn, (int)n_pad, &one, A, CUDA_R_16BF, (int)n_pad,
                            B, CUDA_R_16BF, (int)n_pad, &zero, out, CUDA_R_32F, m, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
    };
    for (int64_t k = 0; k < nb; ++k) {
        const int cur = (int)(k % NR);
        const int64_t s = h->starts[k]; const int b = (int)(h->starts[k + 1] - s);
        const int64_t units = (int64_t)b * (n_pad >> 2);
        if (ms) jw_k_unpack_bf16<true><<<(unsigned)ceil_div(units, 256), 256, 0, h->stream>>>(h->d_packed, h->stride_d, n, n_pad, s, b, C[cur], V[cur]);
        else jw_k_unpack_bf16<false><<<(unsigned)ceil_div(units, 256), 256, 0, h->stream>>>(h->d_packed, h->stride_d, n, n_pad, s, b, C[cur], nullptr);
        JW_LAUNCH_CHECK(h);
        for (int d = 0; d <= cross_max && d <= k; ++d) {
            // d = 0: rows = cols = block k ; d >= 1: rows = block k-d, cols = block k
            const int r = (int)((k - d) % NR);
            const int64_t s_r = h->starts[k - d];
            const int b_r = (int)(h->starts[k - d + 1] - s_r);
            float* out = d == 0 ? h->d_gram + h->gram_off[k] : h->d_gramx[d - 1] + h->gramx_off[d - 1][k];
            JW_CUBLAS(gemm(C[cur], b, C[r], b_r, cnt[0]));                       // Nab[a][c]
            if (ms) {
                JW_CUBLAS(gemm(V[cur], b, C[r], b_r, cnt[1]));                   // sum_i C_r[i,a] V_c[i,c]
                JW_CUBLAS(gemm(C[cur], b, V[r], b_r, cnt[2]));                   // sum_i V_r[i,a] C_c[i,c]
                JW_CUBLAS(gemm(V[cur], b, V[r], b_r, cnt[3]));
            }
            if (N) {
                JW_NCCL(N->GroupStart());
                for (int q = 0; q < (ms ? 4 : 1); ++q)
                    JW_NCCL(N->AllReduce(cnt[q], cnt[q], (size_t)b_r * b, JW_NCCL_FLOAT32, JW_NCCL_SUM, h->nccl_comm, h->stream));
                JW_NCCL(N->GroupEnd());
            }
            if (ms) {
                jw_k_gram_finalize<true><<<(unsigned)ceil_div((int64_t)b_r * b, 256), 256, 0, h->stream>>>(
                    cnt[0], cnt[1], cnt[2], cnt[3], h->n, h->d_means, h->d_colsum, s_r, b_r, s, b, out);
            } else {
                jw_k_gram_finalize<false><<<(unsigned)ceil_div((int64_t)b_r * b, 256), 256, 0, h->stream>>>(
                    cnt[0], nullptr, nullptr, nullptr, h->n, h->d_means, h->d_colsum, s_r, b_r, s, b, out);
            }
            JW_LAUNCH_CHECK(h);
        }
    }
    JW_CUDA(cudaStreamSynchronize(h->stream));
    for (int q = 0; q < NR; ++q) { cudaFree(C[q]); if (V[q]) cudaFree(V[q]); }
    for (int q = 0; q < 4; ++q) if (cnt[q]) cudaFree(cnt[q]);
    cublasDestroy(cb);
    h->gramx_built = cross_max;
    return 0;
}

// Gram blocks (cross = false) or cross-Gram of consecutive blocks (cross = true): every 64x64 tile
static int build_gram(jwas_handle* h, int dist) {
    JW_REQUIRE(h->world == 1, "the popcount Gram kernel does not support sharded rows (use the default GEMM path, nObs < 2^22)");
    const bool cross = dist > 0;
    const int64_t nb = h->nblocks;
    std::vector<int32_t> tblk, tab; std::vector<int64_t> toff;
    std::vector<int64_t> xoff(nb, 0);
    int64_t total = 0;
    for (int64_t i = dist; i < nb; ++i) {
        const int64_t rb = i - dist, cb = i;
        const int64_t br = h->starts[rb + 1] - h->starts[rb], bc = h->starts[cb + 1] - h->starts[cb];
        const int64_t off = cross ? total : h->gram_off[i];
        if (cross) { xoff[i] = total; total += br * bc; }
        const int nta = (int)ceil_div(br, JW_GT), ntb = (int)ceil_div(bc, JW_GT);
        for (int a = 0; a < nta; ++a) for (int c = 0; c < ntb; ++c) {
            tblk.push_back((int32_t)rb); tblk.push_back((int32_t)cb); tab.push_back(a); tab.push_back(c); toff.push_back(off);
        }
    }
    float* out = h->d_gram;
    if (cross) out = h->d_gramx[dist - 1];       // allocated (with these offsets) by alloc_gramx
    if (toff.empty()) return 0;
    int32_t *d_tb = nullptr, *d_tab = nullptr; int64_t* d_toff = nullptr;
    JW_CUDA(cudaMalloc((void**)&d_tb, tblk.size() * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&d_tab, tab.size() * sizeof(int32_t)));
    JW_CUDA(cudaMalloc((void**)&d_toff, toff.size() * sizeof(int64_t)));
    JW_CUDA(cudaMemcpyAsync(d_tb, tblk.data(), tblk.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaMemcpyAsync(d_toff, toff.data(), toff.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    if (h->has_missing)
        jw_k_gram<true><<<(unsigned)toff.size(), 256, 0, h->stream>>>(h->d_packed, h->stride_d, h->n, h->d_means,
            h->d_colsum, h->d_starts, d_toff, d_tb, d_tab, out);
    else
        jw_k_gram<false><<<(unsigned)toff.size(), 256, 0, h->stream>>>(h->d_packed, h->stride_d, h->n, h->d_means,
            h->d_colsum, h->d_starts, d_toff, d_tb, d_tab, out);
    JW_LAUNCH_CHECK(h);
    JW_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_tb); cudaFree(d_tab); cudaFree(d_toff);
    return 0;
}

// the blocks' own Gram and the cross-Gram towards the `lag` previous blocks
static int build_all_gram(jwas_handle* h, int lag) {
    const bool use_gemm = !h->opt_gram_popc && h->n < ((int64_t)1 << 22);   // 4n < 2^24: FP32 sums exact
    if (use_gemm) return build_gram_gemm(h, lag);
    if (alloc_gramx(h, lag)) return 10;
    for (int d = 0; d <= lag; ++d) { int rc = build_gram(h, d); if (rc) return rc; }
    h->gramx_built = lag;
    return 0;
}

extern "C" int jwas_set_blocks(jwas_handle* h, const int64_t* starts, int64_t nblocks) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->stats_ready, "jwas_set_blocks: a row shard needs jwas_init_sharding first (marker statistics are summed over the ranks)");
    JW_REQUIRE(starts && nblocks > 0, "fast_blocks block start vector cannot be empty.");
    JW_REQUIRE(starts[0] == 0, "fast_blocks block starts must begin with 1.");
    JW_REQUIRE(starts[nblocks] == h->p, "fast_blocks block boundaries must end at nMarkers.");
    int64_t maxb = 0, total = 0;
    std::vector<int64_t> off(nblocks);
    for (int64_t i = 0; i < nblocks; ++i) {
        int64_t b = starts[i + 1] - starts[i];
        JW_REQUIRE(b > 0, "fast_blocks block starts must be sorted and unique.");
        JW_REQUIRE(starts[i] >= 0 && starts[i] < h->p, "fast_blocks block starts must be within 1:nMarkers.");
        JW_REQUIRE(b <= JW_MAX_PANEL, "fast_blocks: block size above 4096 is not supported by the GPU backend.");
        off[i] = total; total += b * b; maxb = std::max(maxb, b);
    }
    JW_CUDA(cudaSetDevice(h->device));
    for (void* q : {(void*)h->d_starts, (void*)h->d_gram_off, (void*)h->d_gram}) if (q) cudaFree(q);
    h->d_starts = nullptr; h->d_gram_off = nullptr; h->d_gram = nullptr;
    h->starts.assign(starts, starts + nblocks + 1);
    h->gram_off = off; h->nblocks = nblocks; h->maxb = maxb;
    JW_CUDA(cudaMalloc((void**)&h->d_starts, (nblocks + 1) * sizeof(int64_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_gram_off, nblocks * sizeof(int64_t)));
    JW_CUDA(cudaMalloc((void**)&h->d_gram, (size_t)total * sizeof(float)));
    JW_CUDA(cudaMemcpyAsync(h->d_starts, starts, (nblocks + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaMemcpyAsync(h->d_gram_off, off.data(), nblocks * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    int rc0 = build_all_gram(h, (int)h->opt_lag);
    if (rc0) return rc0;
    int rc = jw_fused_prepare(h);
    if (rc) return rc;
    return 0;
}

// ------------------------------------------------------------------------------------------
// state movement
// ------------------------------------------------------------------------------------------
extern "C" int jwas_put_ycorr(jwas_handle* h, const float* y) {
    JW_REQUIRE(h && y, "jwas_put_ycorr: null argument");
    JW_CUDA(cudaSetDevice(h->device));
    JW_CUDA(cudaMemcpyAsync(h->d_ycorr, y, (size_t)h->t * h->n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    h->next_maxabs = -1.0f;
    return 0;
}
extern "C" int jwas_get_ycorr(jwas_handle* h, float* y) {
    JW_REQUIRE(h && y, "jwas_get_ycorr: null argument");
    JW_CUDA(cudaSetDevice(h->device));
    JW_CUDA(cudaMemcpyAsync(y, h->d_ycorr, (size_t)h->t * h->n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int jwas_put_state(jwas_handle* h, const float* alpha, const float* beta, const int32_t* delta) {
    JW_REQUIRE(h, "null handle");
    JW_CUDA(cudaSetDevice(h->device));
    size_t tp = (size_t)h->t * h->p;
    if (alpha) JW_CUDA(cudaMemcpyAsync(h->d_alpha, alpha, tp * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (beta) JW_CUDA(cudaMemcpyAsync(h->d_beta, beta, tp * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (delta) JW_CUDA(cudaMemcpyAsync(h->d_delta, delta, tp * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int jwas_get_state(jwas_handle* h, float* alpha, float* beta, int32_t* delta) {
    JW_REQUIRE(h, "null handle");
    JW_CUDA(cudaSetDevice(h->device));
    size_t tp = (size_t)h->t * h->p;
    if (alpha) JW_CUDA(cudaMemcpyAsync(alpha, h->d_alpha, tp * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (beta) JW_CUDA(cudaMemcpyAsync(beta, h->d_beta, tp * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (delta) JW_CUDA(cudaMemcpyAsync(delta, h->d_delta, tp * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// multi-GPU: every rank holds (or has just updated) only its own rows [row_begin,row_end) of the t vectors at
// `y` (pitch n); one NCCL all-gather makes them whole on every rank
static int gather_rows(jwas_handle* h, float* y) {
    if (h->world == 1) return 0;
    jw_nccl_api* N = jw_nccl();
    JW_REQUIRE(N && h->nccl_comm, "multi-GPU call without an initialised NCCL communicator");
    const int t = h->t, W = h->world;
    int64_t chunk = 0;
    for (int r = 0; r < W; ++r) chunk = std::max(chunk, h->shard_bounds[r + 1] - h->shard_bounds[r]);
    const size_t need = (size_t)(W + 1) * t * chunk;
    if (ensure_cap(&h->d_gath, &h->cap_gath, need)) return 10;
    float* send = h->d_gath; float* recv = h->d_gath + (size_t)t * chunk;
    for (int k = 0; k < t; ++k)
        JW_CUDA(cudaMemcpyAsync(send + (size_t)k * chunk, y + (size_t)k * h->n + h->row_begin,
                                (size_t)(h->row_end - h->row_begin) * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    JW_NCCL(N->AllGather(send, recv, (size_t)t * chunk, JW_NCCL_FLOAT32, h->nccl_comm, h->stream));
    jw_bounds B;
    for (int r = 0; r <= W; ++r) B.b[r] = h->shard_bounds[r];
    dim3 grid((unsigned)ceil_div(chunk, 256), (unsigned)(W * t));
    jw_k_scatter_rows<<<grid, 256, 0, h->stream>>>(recv, chunk, t, W, B, y, h->n, h->rank);
    JW_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int jwas_ycorr_sub_malpha(jwas_handle* h) {
    JW_REQUIRE(h, "null handle");
    JW_CUDA(cudaSetDevice(h->device));
    for (int k = 0; k < h->t; ++k) {
        jw_k_mul_alpha<<<(unsigned)ceil_div(jw_nloc(h), 256), 256, 0, h->stream>>>(
            jw_packed_g(h), h->stride_d, h->row_begin, h->row_end, h->p, h->d_means, h->d_alpha + (size_t)k * h->p, -1.0f,
            h->d_ycorr + (size_t)k * h->n, 1);
        JW_LAUNCH_CHECK(h);
    }
    if (gather_rows(h, h->d_ycorr)) return 13;
    JW_CUDA(cudaStreamSynchronize(h->stream));
    h->next_maxabs = -1.0f;
    return 0;
}
extern "C" int jwas_mul_alpha(jwas_handle* h, int trait, float* out) {
    JW_REQUIRE(h && out, "jwas_mul_alpha: null argument");
    JW_REQUIRE(trait >= 0 && trait < h->t, "jwas_mul_alpha: trait out of range");
    JW_CUDA(cudaSetDevice(h->device));
    float* d_out = (float*)h->d_yq + (size_t)trait * h->n;   // scratch (re-quantised by the next sweep anyway)
    jw_k_mul_alpha<<<(unsigned)ceil_div(jw_nloc(h), 256), 256, 0, h->stream>>>(
        jw_packed_g(h), h->stride_d, h->row_begin, h->row_end, h->p, h->d_means, h->d_alpha + (size_t)trait * h->p, 1.0f, d_out, 0);
    JW_LAUNCH_CHECK(h);
    if (h->world > 1) {                 // the other ranks' rows (gather_rows moves all t vectors of the scratch)
        if (gather_rows(h, (float*)h->d_yq)) return 13;
    }
    JW_CUDA(cudaMemcpyAsync(out, d_out, h->n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// canonical chunked sum of a[i]*b[i] -> *d_out (device)
static int canonical_dot(jwas_handle* h, const float* a, const float* b, int64_t n, double* d_out) {
    int64_t nchunks = ceil_div(n, JW_CHUNK);
    JW_REQUIRE(nchunks <= (int64_t)JW_CHUNK * JW_CHUNK, "vector too long for the canonical reduction (16.7M elements)");
    if (ensure_cap(&h->d_partials, &h->cap_partials, (size_t)nchunks)) return 10;
    jw_k_chunk_prod<<<(unsigned)ceil_div(nchunks, 128), 128, 0, h->stream>>>(a, b, n, h->d_partials);
    JW_LAUNCH_CHECK(h);
    jw_k_chunk_final<<<1, JW_CHUNK, 0, h->stream>>>(h->d_partials, nchunks, d_out);
    JW_LAUNCH_CHECK(h);
    return 0;
}

static int canonical_sum(jwas_handle* h, const float* a, int64_t n, double* d_out) {
    int64_t nchunks = ceil_div(n, JW_CHUNK);
    JW_REQUIRE(nchunks <= (int64_t)JW_CHUNK * JW_CHUNK, "vector too long for the canonical reduction (16.7M elements)");
    if (ensure_cap(&h->d_partials, &h->cap_partials, (size_t)nchunks)) return 10;
    jw_k_chunk_sum<<<(unsigned)ceil_div(nchunks, 128), 128, 0, h->stream>>>(a, n, h->d_partials);
    JW_LAUNCH_CHECK(h);
    jw_k_chunk_final<<<1, JW_CHUNK, 0, h->stream>>>(h->d_partials, nchunks, d_out);
    JW_LAUNCH_CHECK(h);
    return 0;
}

extern "C" int jwas_shift_ycorr(jwas_handle* h, int trait, float shift, double* sum, double* sumsq) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(trait >= 0 && trait < h->t, "jwas_shift_ycorr: trait out of range");
    JW_CUDA(cudaSetDevice(h->device));
    float* y = h->d_ycorr + (size_t)trait * h->n;
    if (shift != 0.0f) {
        jw_k_shift<<<(unsigned)ceil_div(h->n, 256), 256, 0, h->stream>>>(y, h->n, shift);
        JW_LAUNCH_CHECK(h);
        h->next_maxabs = -1.0f;
    }
    double host[2] = {0, 0};
    if (!sum && !sumsq) { JW_CUDA(cudaStreamSynchronize(h->stream)); return 0; }
    if (sumsq) { if (canonical_dot(h, y, y, h->n, h->d_stats)) return 10; }
    if (sum) { if (canonical_sum(h, y, h->n, h->d_stats + 1)) return 10; }
    JW_CUDA(cudaMemcpyAsync(host, h->d_stats, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    if (sumsq) *sumsq = host[0];
    if (sum) *sum = host[1];
    return 0;
}

// ------------------------------------------------------------------------------------------
// sweep driver
// ------------------------------------------------------------------------------------------
struct sweep_cfg {
    int method = 0, schedule = 0, full_reps = 1;
    double vare = 1.0, sigmaSq = 0.0;
    int nclasses = 0, per_marker_pi = 0, per_marker_G = 0;
    double gamma[JW_MAX_CLASSES] = {0};
    double Rinv[16] = {0}, Ginv[16] = {0};
    double pi_host[16] = {0};     // global class / joint-state priors (BayesR, multi-trait)
    int mega = 0;
    uint64_t seed = 0; uint32_t iter = 0;
    const double* u = nullptr; const double* z = nullptr;  // device pointers
};

template <int METHOD, int T>
static void launch_chain(jwas_handle* h, const jw_chain_args& A, int nblk, int threads) {
    const int list_cap = h->maxb > JW_MAX_BLOCK ? (int)h->maxb : 0;
    const size_t smem = jw_chain_smem_bytes(T, list_cap);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(jw_k_chain<METHOD, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    jw_k_chain<METHOD, T><<<nblk, threads, smem, h->stream>>>(A, list_cap);
}
static int dispatch_chain(jwas_handle* h, const jw_chain_args& A, int nblk, int threads) {
    int t = h->t;
    if (A.method == 0 && t == 1) launch_chain<0, 1>(h, A, nblk, threads);
    else if (A.method == 1 && t == 1) launch_chain<1, 1>(h, A, nblk, threads);
    else if (A.method == 2 && t == 2) launch_chain<2, 2>(h, A, nblk, threads);
    else if (A.method == 2 && t == 3) launch_chain<2, 3>(h, A, nblk, threads);
    else if (A.method == 2 && t == 4) launch_chain<2, 4>(h, A, nblk, threads);
    else if (A.method == 3 && t == 2) launch_chain<3, 2>(h, A, nblk, threads);
    else if (A.method == 4 && t == 2) launch_chain<4, 2>(h, A, nblk, threads);
    else if (A.method == 4 && t == 3) launch_chain<4, 3>(h, A, nblk, threads);
    else if (A.method == 4 && t == 4) launch_chain<4, 4>(h, A, nblk, threads);
    else { jw_set_error("unsupported (method, traits) combination"); return 2; }
    JW_LAUNCH_CHECK(h);
    return 0;
}
static int prof_begin(jwas_handle* h) {
    if (!h->opt_profile) return 0;
    cudaEvent_t a, b;
    JW_CUDA(cudaEventCreate(&a)); JW_CUDA(cudaEventCreate(&b));
    h->prof_events.push_back(a); h->prof_events.push_back(b);
    JW_CUDA(cudaEventRecord(a, h->stream));
    return 0;
}
static int prof_end(jwas_handle* h) {
    if (!h->opt_profile) return 0;
    JW_CUDA(cudaEventRecord(h->prof_events.back(), h->stream));
    return 0;
}
static void prof_collect(jwas_handle* h) {
    h->prof_ms = 0.0; h->prof_launches = (int64_t)h->prof_events.size() / 2;
    for (size_t i = 0; i + 1 < h->prof_events.size(); i += 2) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, h->prof_events[i], h->prof_events[i + 1]);
        h->prof_ms += ms;
        cudaEventDestroy(h->prof_events[i]); cudaEventDestroy(h->prof_events[i + 1]);
    }
    h->prof_events.clear();
}
static int dispatch_dot(jwas_handle* h, int64_t j0, int64_t nj) {
    if (prof_begin(h)) return 10;
    const int64_t w0 = h->row_begin >> 4, w1 = ceil_div(h->row_end, 16);
    dim3 grid((unsigned)ceil_div(nj, 8), (unsigned)std::max<int64_t>(1, ceil_div(w1 - w0, JW_DOT_SLAB_WORDS)));
    long long* mq = h->d_mq;
#define JW_DOT(T_, M_) jw_k_block_dot<T_, M_><<<grid, 256, 0, h->stream>>>(jw_packed_g(h), h->stride_d, h->n, h->p, j0, nj, h->d_yq, h->d_dq, mq, w0, w1)
    bool ms = h->has_missing != 0;
    switch (h->t) {
        case 1: if (ms) JW_DOT(1, true); else JW_DOT(1, false); break;
        case 2: if (ms) JW_DOT(2, true); else JW_DOT(2, false); break;
        case 3: if (ms) JW_DOT(3, true); else JW_DOT(3, false); break;
        default: if (ms) JW_DOT(4, true); else JW_DOT(4, false); break;
    }
#undef JW_DOT
    JW_LAUNCH_CHECK(h);
    if (prof_end(h)) return 10;
    return 0;
}
static int dispatch_apply(jwas_handle* h) {
    unsigned g = (unsigned)std::max<int64_t>(1, ceil_div(h->row_end - h->row_begin, 256));
#define JW_APPLY(T_) jw_k_apply<T_><<<g, 256, 0, h->stream>>>(jw_packed_g(h), h->stride_d, h->n, h->p, h->d_means, h->d_dalpha, h->d_act_idx, h->d_act_cnt, h->d_ycorr, h->row_begin, h->row_end)
    switch (h->t) { case 1: JW_APPLY(1); break; case 2: JW_APPLY(2); break; case 3: JW_APPLY(3); break; default: JW_APPLY(4); }
#undef JW_APPLY
    JW_LAUNCH_CHECK(h);
    return 0;
}

static int collect_stats(jwas_handle* h, const sweep_cfg& c, int S, jwas_sweep_stats* st) {
    const int t = h->t;
    memset(st, 0, sizeof(*st));
    // d_stats layout: [0..15] ycorr_ss, [16..31] alpha_ss, [32..47] beta_ss, [48] bayesr ssq
    for (int a = 0; a < t; ++a)
        for (int b = a; b < t; ++b) {
            if (canonical_dot(h, h->d_ycorr + (size_t)a * h->n, h->d_ycorr + (size_t)b * h->n, h->n, h->d_stats + a * t + b)) return 10;
            if (canonical_dot(h, h->d_alpha + (size_t)a * h->p, h->d_alpha + (size_t)b * h->p, h->p, h->d_stats + 16 + a * t + b)) return 10;
            if (c.method != 1)
                if (canonical_dot(h, h->d_beta + (size_t)a * h->p, h->d_beta + (size_t)b * h->p, h->p, h->d_stats + 32 + a * t + b)) return 10;
        }
    for (int a = 0; a < t; ++a)
        if (canonical_sum(h, h->d_ycorr + (size_t)a * h->n, h->n, h->d_stats + 49 + a)) return 10;
    if (c.method == 1) {
        int64_t nchunks = ceil_div(h->p, JW_CHUNK);
        if (ensure_cap(&h->d_partials, &h->cap_partials, (size_t)nchunks)) return 10;
        jw_k_chunk_bayesr<<<(unsigned)ceil_div(nchunks, 128), 128, 0, h->stream>>>(h->d_alpha, h->d_delta, h->p,
            c.gamma[1], c.gamma[2], c.gamma[3], c.gamma[4], c.gamma[5], c.gamma[6], c.gamma[7], h->d_partials);
        JW_LAUNCH_CHECK(h);
        jw_k_chunk_final<<<1, JW_CHUNK, 0, h->stream>>>(h->d_partials, nchunks, h->d_stats + 48);
        JW_LAUNCH_CHECK(h);
    }
    JW_CUDA(cudaMemsetAsync(h->d_counters + 4, 0, 24 * sizeof(unsigned long long), h->stream));
    jw_k_counts<<<(unsigned)ceil_div(h->p, 256), 256, 0, h->stream>>>(h->d_alpha, h->d_delta, h->p, t, c.method, h->d_counters + 4);
    JW_LAUNCH_CHECK(h);
    JW_CUDA(cudaMemsetAsync(h->d_maxabs, 0, sizeof(float), h->stream));
    jw_k_maxabs<<<64, 256, 0, h->stream>>>(h->d_ycorr, (int64_t)t * h->n, (unsigned*)h->d_maxabs);
    JW_LAUNCH_CHECK(h);

    double hs[53]; unsigned long long hc[28]; int32_t hf[4]; float hm;
    JW_CUDA(cudaMemcpyAsync(hs, h->d_stats, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaMemcpyAsync(hc, h->d_counters, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaMemcpyAsync(hf, h->d_flags, sizeof(hf), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaMemcpyAsync(&hm, h->d_maxabs, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaEventRecord(h->ev1, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->last_sweep_ms = ms;
    prof_collect(h);
    for (int a = 0; a < t; ++a)
        for (int b = a; b < t; ++b) {
            st->ycorr_ss[a * t + b] = st->ycorr_ss[b * t + a] = hs[a * t + b];
            st->alpha_ss[a * t + b] = st->alpha_ss[b * t + a] = hs[16 + a * t + b];
            st->beta_ss[a * t + b] = st->beta_ss[b * t + a] = hs[32 + a * t + b];
        }
    st->bayesr_ssq = hs[48];
    for (int a = 0; a < t; ++a) st->ycorr_sum[a] = hs[49 + a];
    st->n_active = (int64_t)hc[0]; st->n_rounds = (int64_t)hc[1];
    for (int k = 0; k < t; ++k) { st->nnz_alpha[k] = (double)hc[4 + k]; st->sum_delta[k] = (double)hc[8 + k]; }
    for (int q = 0; q < 16; ++q) st->class_counts[q] = (double)hc[12 + q];
    st->ycorr_maxabs = hm; st->scale_exp = S; st->overflow = hf[0];
    h->next_maxabs = hm;
    if (hf[0]) { jw_set_error("ycorr fixed-point overflow: |ycorr| grew more than 4x inside one sweep"); return 3; }
    return 0;
}

// multi-GPU: exact int64 partial rhs of markers [j0, j0+nj) summed over ranks (C1 of SURVEY 2a)
static int reduce_block_rhs(jwas_handle* h, int64_t j0, int64_t nj) {
    if (h->world == 1) return 0;
    jw_nccl_api* N = jw_nccl();
    JW_REQUIRE(N && h->nccl_comm, "multi-GPU sweep without an initialised NCCL communicator");
    JW_NCCL(N->GroupStart());
    for (int k = 0; k < h->t; ++k) {
        long long* dq = h->d_dq + (size_t)k * h->p + j0;
        JW_NCCL(N->AllReduce(dq, dq, (size_t)nj, JW_NCCL_INT64, JW_NCCL_SUM, h->nccl_comm, h->stream));
        if (h->has_missing) {
            long long* mq = h->d_mq + (size_t)k * h->p + j0;
            JW_NCCL(N->AllReduce(mq, mq, (size_t)nj, JW_NCCL_INT64, JW_NCCL_SUM, h->nccl_comm, h->stream));
        }
    }
    JW_NCCL(N->AllReduce(h->d_sq, h->d_sq, (size_t)h->t, JW_NCCL_INT64, JW_NCCL_SUM, h->nccl_comm, h->stream));
    JW_NCCL(N->GroupEnd());
    return 0;
}
static int gather_ycorr(jwas_handle* h) { return gather_rows(h, h->d_ycorr); }

static int run_sweep(jwas_handle* h, sweep_cfg& c, jwas_sweep_stats* st) {
    JW_REQUIRE(h->nblocks > 0, "jwas_set_blocks must be called before a sweep");
    JW_REQUIRE(st != nullptr, "stats is NULL");
    const int t = h->t;
    const int64_t n = h->n, p = h->p;
    JW_CUDA(cudaEventRecord(h->ev0, h->stream));

    // fixed-point scale from max|ycorr| (carried over from the previous sweep when untouched)
    float maxabs = h->next_maxabs;
    if (maxabs < 0.0f) {
        JW_CUDA(cudaMemsetAsync(h->d_maxabs, 0, sizeof(float), h->stream));
        jw_k_maxabs<<<64, 256, 0, h->stream>>>(h->d_ycorr, (int64_t)t * n, (unsigned*)h->d_maxabs);
        JW_LAUNCH_CHECK(h);
        JW_CUDA(cudaMemcpyAsync(&maxabs, h->d_maxabs, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        JW_CUDA(cudaStreamSynchronize(h->stream));
    }
    const int S = jw_choose_scale_exp(maxabs);
    const float scale = jw_pow2f(S);
    const double invscale = (double)jw_pow2f(-S);

    JW_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t), h->stream));
    JW_CUDA(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(unsigned long long), h->stream));
    JW_CUDA(cudaMemsetAsync(h->d_counters + 32, 0, 32 * sizeof(unsigned long long), h->stream));    // phase timers

    jw_chain_args A;
    memset(&A, 0, sizeof(A));
    A.n = n; A.p = p; A.t = t; A.method = c.method; A.nclasses = c.nclasses;
    A.nreps_mode = (c.schedule == JWAS_SCHED_EXACT) ? 0 : c.full_reps;
    A.per_marker_pi = c.per_marker_pi; A.per_marker_G = c.per_marker_G;
    A.starts = h->d_starts; A.gram_off = h->d_gram_off; A.gram = h->d_gram;
    A.means = h->d_means; A.xpx = h->d_xpx;
    A.dq = h->d_dq; A.mq = h->has_missing ? h->d_mq : nullptr; A.sq = h->d_sq;
    A.invscale = invscale;
    A.alpha = h->d_alpha; A.beta = h->d_beta; A.delta = h->d_delta; A.dalpha = h->d_dalpha;
    A.vare = c.vare; A.sigmaSq = c.sigmaSq;
    A.ve = h->d_ve; A.pi = h->d_pi; A.bigPi = h->d_pi; A.Gmat = h->d_ve;
    memcpy(A.gamma, c.gamma, sizeof(A.gamma));
    memcpy(A.Rinv, c.Rinv, sizeof(A.Rinv)); memcpy(A.Ginv, c.Ginv, sizeof(A.Ginv));
    A.seed = c.seed; A.iter = c.iter; A.u = c.u; A.z = c.z;
    if (c.mega) for (int k = 0; k < t; ++k) A.lpi[k] = c.pi_host[k];     // mega: the per-trait pi itself
    A.act_idx = h->d_act_idx; A.act_cnt = h->d_act_cnt; A.counters = h->d_counters; A.timers = (int)h->opt_timers;
    const int threads = (int)std::min<int64_t>(JW_MAX_BLOCK, std::max<int64_t>(32, ceil_div(h->maxb, 32) * 32));
    JW_REQUIRE(c.schedule == JWAS_SCHED_EXACT || h->maxb <= JW_MAX_BLOCK,
               "fast_blocks: block size above 1024 is supported by the exact schedule only.");
    if (c.method == 0) {
        // draw-independent chain terms for every marker, computed by the whole GPU up front
        if (!h->d_prep) {
            JW_CUDA(cudaMalloc((void**)&h->d_prep, (size_t)6 * p * sizeof(double)));
            JW_CUDA(cudaMalloc((void**)&h->d_prep_beta0, (size_t)p * sizeof(float)));
        }
        jw_k_prep_abc<<<(unsigned)ceil_div(p, 256), 256, 0, h->stream>>>(A, h->d_prep, h->d_prep_beta0);
        JW_LAUNCH_CHECK(h);
        A.prep = h->d_prep; A.prep_beta0 = h->d_prep_beta0;
    } else {
        // BayesR / multi-trait: the draws of repetition 0 (Philox, Box-Muller, log-odds thresholds)
        if (ensure_cap(&h->d_draws, &h->cap_draws, (size_t)2 * t * p)) return 10;
        jw_k_prep_draws<<<(unsigned)ceil_div((int64_t)t * p, 256), 256, 0, h->stream>>>(A, h->d_draws, h->d_draws + (size_t)t * p);
        JW_LAUNCH_CHECK(h);
        A.draws_u = h->d_draws; A.draws_z = h->d_draws + (size_t)t * p;
        // rhs-independent terms per marker on all SMs; logs of the global priors once on the host
        if (!c.per_marker_pi && !c.per_marker_G) {
            const size_t cnt = (size_t)(c.method == 1 ? 2 * (c.nclasses - 1) : t) * p;
            if (ensure_cap(&h->d_prep_rm, &h->cap_prep_rm, cnt)) return 10;
            A.host_logs = 1;
            if (c.method == 1) for (int k = 0; k < c.nclasses; ++k) A.lpi[k] = jw_log(c.pi_host[k]);
            else {
                for (int k = 0; k < t; ++k) A.mt_lG[k] = jw_log(c.Ginv[k * t + k]);
                for (int q = 0; q < (1 << t); ++q) A.mt_lPi[q] = jw_log(c.pi_host[q]);
            }
            jw_k_prep_rm<<<(unsigned)ceil_div(p, 256), 256, 0, h->stream>>>(A, h->d_prep_rm);
            JW_LAUNCH_CHECK(h);
            A.prep_rm = h->d_prep_rm;
        }
    }

    // engine 1 (persistent fused kernel) where it covers the shape; engine 0 otherwise.  The lagged schedule is
    // the exact schedule's; block repetitions (nreps = block size) run engine 1 without the lag.
    const bool fused_ok = h->fused != nullptr && ((jw_fused_state*)h->fused)->ready;
    const bool fused_multi_ok = h->world == 1 || (h->ipc_ready && h->opt_lag && A.nreps_mode == 0);
    if (h->opt_engine == 1 && fused_ok && c.schedule != JWAS_SCHED_INDEPENDENT && fused_multi_ok &&
        (A.nreps_mode == 0 || ((jw_fused_state*)h->fused)->legacy_ok)) {
        int rc = jw_fused_sweep(h, A, scale);
        if (rc) return rc;
        if (gather_ycorr(h)) return 13;
        return collect_stats(h, c, S, st);
    }

    JW_CUDA(cudaMemsetAsync(h->d_dq, 0, (size_t)t * p * sizeof(long long), h->stream));
    if (h->has_missing) JW_CUDA(cudaMemsetAsync(h->d_mq, 0, (size_t)t * p * sizeof(long long), h->stream));
    dim3 qgrid((unsigned)std::min<int64_t>(ceil_div(n, 256), 1024), (unsigned)t);

    if (c.schedule == JWAS_SCHED_INDEPENDENT) {
        // all blocks read the entry snapshot (BayesABC.jl:205); one GEMV over all of M
        JW_CUDA(cudaMemsetAsync(h->d_sq, 0, JW_MAX_TRAITS * sizeof(long long), h->stream));
        if (h->opt_engine == 1 && jw_stream_supported(h)) {
            // engine 1: the rhs of every block in ONE streamed pass over the tiled image (lookup-table stream,
            // tables built once per row slice, no per-panel barrier)
            if (prof_begin(h)) return 10;
            int rcs = jw_stream_all(h, scale);
            if (rcs) return rcs;
            if (prof_end(h)) return 10;
        } else {
            jw_k_quantize<<<qgrid, 256, 0, h->stream>>>(h->d_ycorr, n, t, scale, h->d_yq, h->d_sq, h->d_flags, h->row_begin, h->row_end);
            JW_LAUNCH_CHECK(h);
            if (dispatch_dot(h, 0, p)) return 11;
        }
        if (reduce_block_rhs(h, 0, p)) return 13;
        A.block0 = 0; A.write_active_list = 0;
        if (dispatch_chain(h, A, (int)h->nblocks, threads)) return 11;
        jw_k_compact_active<<<1, 1024, 0, h->stream>>>(h->d_dalpha, p, t, h->d_act_idx, h->d_act_cnt);
        JW_LAUNCH_CHECK(h);
        if (dispatch_apply(h)) return 11;
    } else {
        for (int64_t ib = 0; ib < h->nblocks; ++ib) {
            JW_CUDA(cudaMemsetAsync(h->d_sq, 0, JW_MAX_TRAITS * sizeof(long long), h->stream));
            jw_k_quantize<<<qgrid, 256, 0, h->stream>>>(h->d_ycorr, n, t, scale, h->d_yq, h->d_sq, h->d_flags, h->row_begin, h->row_end);
            JW_LAUNCH_CHECK(h);
            if (dispatch_dot(h, h->starts[ib], h->starts[ib + 1] - h->starts[ib])) return 11;
            if (reduce_block_rhs(h, h->starts[ib], h->starts[ib + 1] - h->starts[ib])) return 13;
            A.block0 = (int)ib; A.write_active_list = 1;
            if (dispatch_chain(h, A, 1, threads)) return 11;
            if (dispatch_apply(h)) return 11;
        }
    }
    if (gather_ycorr(h)) return 13;
    return collect_stats(h, c, S, st);
}

static int upload_doubles(jwas_handle* h, double** dptr, size_t* cap, const double* src, size_t count) {
    if (ensure_cap(dptr, cap, count)) return 10;
    JW_CUDA(cudaMemcpyAsync(*dptr, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return 0;
}
static int fill_doubles(jwas_handle* h, double** dptr, size_t* cap, double v, size_t count) {
    if (ensure_cap(dptr, cap, count)) return 10;
    jw_k_fill_double<<<(unsigned)ceil_div((int64_t)count, 256), 256, 0, h->stream>>>(*dptr, (int64_t)count, v);
    JW_LAUNCH_CHECK(h);
    return 0;
}
static int upload_draws(jwas_handle* h, sweep_cfg& c, const double* u, const double* z, int schedule, int full_reps) {
    c.u = nullptr; c.z = nullptr;
    if (!u && !z) return 0;
    JW_REQUIRE(u && z, "replay draws: u and z must both be given");
    size_t reps = (schedule == JWAS_SCHED_EXACT || !full_reps) ? 1 : (size_t)h->maxb;
    size_t count = reps * h->t * h->p;
    if (upload_doubles(h, &h->d_u, &h->cap_u, u, count)) return 10;
    if (upload_doubles(h, &h->d_z, &h->cap_z, z, count)) return 10;
    c.u = h->d_u; c.z = h->d_z;
    return 0;
}

extern "C" int jwas_sweep_bayesabc(jwas_handle* h, int schedule, double vare, const double* var_effects,
                                   const double* pi, uint64_t seed, uint32_t iter, const double* u,
                                   const double* z, jwas_sweep_stats* stats) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->t == 1, "jwas_sweep_bayesabc: single-trait handle required");
    JW_REQUIRE(schedule >= 0 && schedule <= 2, "unknown schedule");
    JW_REQUIRE(vare > 0.0, "residual variance must be positive");
    JW_CUDA(cudaSetDevice(h->device));
    if (var_effects) { if (upload_doubles(h, &h->d_ve, &h->cap_ve, var_effects, (size_t)h->p)) return 10; }
    else JW_REQUIRE(h->d_ve && h->cap_ve >= (size_t)h->p, "var_effects is NULL and no device-resident vector exists");
    if (pi) { if (upload_doubles(h, &h->d_pi, &h->cap_pi, pi, (size_t)h->p)) return 10; }
    else JW_REQUIRE(h->d_pi && h->cap_pi >= (size_t)h->p, "pi is NULL and no device-resident vector exists");
    sweep_cfg c; c.method = 0; c.schedule = schedule; c.full_reps = 1; c.vare = vare; c.seed = seed; c.iter = iter;
    if (upload_draws(h, c, u, z, schedule, 1)) return 10;
    return run_sweep(h, c, stats);
}

extern "C" int jwas_sweep_bayesc(jwas_handle* h, int schedule, double vare, double var_effect, double pi,
                                 uint64_t seed, uint32_t iter, jwas_sweep_stats* stats) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->t == 1, "jwas_sweep_bayesc: single-trait handle required");
    JW_REQUIRE(schedule >= 0 && schedule <= 2, "unknown schedule");
    JW_REQUIRE(vare > 0.0 && var_effect > 0.0, "variances must be positive");
    JW_REQUIRE(pi >= 0.0 && pi <= 1.0, "pi must lie in [0,1]");
    JW_CUDA(cudaSetDevice(h->device));
    if (fill_doubles(h, &h->d_ve, &h->cap_ve, var_effect, (size_t)h->p)) return 10;
    if (fill_doubles(h, &h->d_pi, &h->cap_pi, pi, (size_t)h->p)) return 10;
    sweep_cfg c; c.method = 0; c.schedule = schedule; c.full_reps = 1; c.vare = vare; c.seed = seed; c.iter = iter;
    return run_sweep(h, c, stats);
}

// BayesABC!(xArray, xRinvArray, xpRinvx, yCorr, alpha, beta, delta, vare, varEffects, pi) mutates the caller's
// arrays in place (BayesABC.jl:60-63).  This is that call for host arrays: copies in, the sweep, copies out, enqueued
// back to back on the handle's stream with ONE synchronisation at the end (pinned buffers copy asynchronously).
extern "C" int jwas_sweep_bayesc_host(jwas_handle* h, int schedule, double vare, double var_effect, double pi,
                                      uint64_t seed, uint32_t iter, float* ycorr, float* alpha, float* beta,
                                      int32_t* delta, jwas_sweep_stats* stats) {
    JW_REQUIRE(h && ycorr && alpha && beta && delta, "jwas_sweep_bayesc_host: null argument");
    JW_REQUIRE(h->t == 1, "jwas_sweep_bayesc_host: single-trait handle required");
    JW_CUDA(cudaSetDevice(h->device));
    const size_t nb = (size_t)h->n * sizeof(float), pb = (size_t)h->p * sizeof(float);
    JW_CUDA(cudaMemcpyAsync(h->d_ycorr, ycorr, nb, cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaMemcpyAsync(h->d_alpha, alpha, pb, cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaMemcpyAsync(h->d_beta, beta, pb, cudaMemcpyHostToDevice, h->stream));
    JW_CUDA(cudaMemcpyAsync(h->d_delta, delta, pb, cudaMemcpyHostToDevice, h->stream));
    h->next_maxabs = -1.0f;
    int rc = jwas_sweep_bayesc(h, schedule, vare, var_effect, pi, seed, iter, stats);
    if (rc) return rc;
    JW_CUDA(cudaMemcpyAsync(ycorr, h->d_ycorr, nb, cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaMemcpyAsync(alpha, h->d_alpha, pb, cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaMemcpyAsync(beta, h->d_beta, pb, cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaMemcpyAsync(delta, h->d_delta, pb, cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int jwas_sweep_bayesr(jwas_handle* h, int schedule, int full_reps, double vare, double sigma_sq,
                                 const double* pi, int per_marker_pi, const double* gamma, int nclasses,
                                 uint64_t seed, uint32_t iter, const double* u, const double* z,
                                 jwas_sweep_stats* stats) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->t == 1, "jwas_sweep_bayesr: single-trait handle required");
    JW_REQUIRE(schedule >= 0 && schedule <= 2, "unknown schedule");
    JW_REQUIRE(nclasses == JW_R_CLASSES, "BayesR Pi must have length 4.");
    JW_REQUIRE(pi && gamma, "BayesR pi/gamma missing");
    JW_REQUIRE(sigma_sq > 0.0, "BayesR sigmaSq must be positive.");
    JW_REQUIRE(vare > 0.0, "residual variance must be positive");
    if (!per_marker_pi) {
        double s = 0.0;
        for (int k = 0; k < nclasses; ++k) { JW_REQUIRE(pi[k] >= 0.0, "BayesR pi entries must be nonnegative."); s += pi[k]; }
        JW_REQUIRE(std::fabs(s - 1.0) <= 1e-8, "BayesR pi must sum to 1.");
    }
    JW_CUDA(cudaSetDevice(h->device));
    size_t npi = per_marker_pi ? (size_t)h->p * nclasses : (size_t)nclasses;
    if (upload_doubles(h, &h->d_pi, &h->cap_pi, pi, npi)) return 10;
    sweep_cfg c; c.method = 1; c.schedule = schedule; c.full_reps = full_reps ? 1 : 0; c.vare = vare;
    c.sigmaSq = sigma_sq; c.nclasses = nclasses; c.per_marker_pi = per_marker_pi;
    for (int k = 0; k < nclasses; ++k) c.gamma[k] = gamma[k];
    if (!per_marker_pi) for (int k = 0; k < nclasses; ++k) c.pi_host[k] = pi[k];
    c.seed = seed; c.iter = iter;
    if (upload_draws(h, c, u, z, schedule, c.full_reps)) return 10;
    return run_sweep(h, c, stats);
}

extern "C" int jwas_sweep_mt1(jwas_handle* h, int schedule, const double* R, const double* G, int per_marker_G,
                              const double* big_pi, int per_marker_pi, uint64_t seed, uint32_t iter,
                              const double* u, const double* z, jwas_sweep_stats* stats) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->t >= 2, "jwas_sweep_mt1: multi-trait handle required");
    JW_REQUIRE(schedule >= 0 && schedule <= 2, "unknown schedule");
    JW_REQUIRE(R && G && big_pi, "jwas_sweep_mt1: R, G and big_pi are required");
    JW_CUDA(cudaSetDevice(h->device));
    const int t = h->t;
    sweep_cfg c; c.method = 2; c.schedule = schedule; c.full_reps = 1; c.seed = seed; c.iter = iter;
    c.per_marker_G = per_marker_G; c.per_marker_pi = per_marker_pi;
    jw_inv_spd_fixed(R, t, c.Rinv);
    if (per_marker_G) { if (upload_doubles(h, &h->d_ve, &h->cap_ve, G, (size_t)h->p * t * t)) return 10; }
    else jw_inv_spd_fixed(G, t, c.Ginv);
    if (!per_marker_pi) for (int q = 0; q < (1 << t); ++q) c.pi_host[q] = big_pi[q];
    size_t npi = (size_t)(per_marker_pi ? h->p : 1) << t;
    if (upload_doubles(h, &h->d_pi, &h->cap_pi, big_pi, npi)) return 10;
    if (upload_draws(h, c, u, z, schedule, 1)) return 10;
    return run_sweep(h, c, stats);
}

extern "C" int jwas_sweep_mt2(jwas_handle* h, int schedule, const double* R, const double* G, const double* big_pi,
                              uint64_t seed, uint32_t iter, const double* u, const double* z, jwas_sweep_stats* stats) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->t == 2, "jwas_sweep_mt2: sampler II is implemented for exactly 2 traits");
    JW_REQUIRE(schedule >= 0 && schedule <= 2, "unknown schedule");
    JW_REQUIRE(R && G && big_pi, "jwas_sweep_mt2: R, G and big_pi are required");
    JW_CUDA(cudaSetDevice(h->device));
    sweep_cfg c; c.method = 3; c.schedule = schedule; c.full_reps = 1; c.seed = seed; c.iter = iter;
    jw_inv_spd_fixed(R, 2, c.Rinv);
    jw_inv_spd_fixed(G, 2, c.Ginv);
    for (int q = 0; q < 4; ++q) c.pi_host[q] = big_pi[q];
    c.per_marker_pi = 1;          // keeps the generic prep of sampler I out of the way (no hoisted logs here)
    if (upload_doubles(h, &h->d_pi, &h->cap_pi, big_pi, 4)) return 10;
    if (upload_draws(h, c, u, z, schedule, 1)) return 10;
    int rc = run_sweep(h, c, stats);
    return rc;
}

extern "C" int jwas_sweep_mega(jwas_handle* h, int schedule, const double* vare, const double* var_effects,
                               const double* pi, uint64_t seed, uint32_t iter, const double* u, const double* z,
                               jwas_sweep_stats* stats) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->t >= 2, "jwas_sweep_mega: multi-trait handle required");
    JW_REQUIRE(schedule >= 0 && schedule <= 2, "unknown schedule");
    JW_REQUIRE(vare && var_effects && pi, "jwas_sweep_mega: vare, var_effects and pi are required");
    JW_CUDA(cudaSetDevice(h->device));
    sweep_cfg c; c.method = 4; c.schedule = schedule; c.full_reps = 1; c.seed = seed; c.iter = iter;
    for (int k = 0; k < h->t; ++k) {
        JW_REQUIRE(vare[k] > 0.0 && var_effects[k] > 0.0, "variances must be positive");
        JW_REQUIRE(pi[k] >= 0.0 && pi[k] <= 1.0, "pi must lie in [0,1]");
        c.Rinv[k] = 1.0 / vare[k]; c.Ginv[k] = var_effects[k]; c.pi_host[k] = pi[k];
    }
    c.per_marker_pi = 1;          // no sampler-I prep
    c.mega = 1;
    if (upload_draws(h, c, u, z, schedule, 1)) return 10;
    return run_sweep(h, c, stats);
}

extern "C" int jwas_fill_hyper(jwas_handle* h, int which, double value) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(which == 0 || which == 1, "jwas_fill_hyper: which must be 0 (var_effects) or 1 (pi)");
    JW_CUDA(cudaSetDevice(h->device));
    if (which == 0) { JW_REQUIRE(value > 0.0, "variances must be positive"); if (fill_doubles(h, &h->d_ve, &h->cap_ve, value, (size_t)h->p)) return 10; }
    else { JW_REQUIRE(value >= 0.0 && value <= 1.0, "pi must lie in [0,1]"); if (fill_doubles(h, &h->d_pi, &h->cap_pi, value, (size_t)h->p)) return 10; }
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int jwas_sample_bayesb_variances(jwas_handle* h, double df, double scale, uint64_t seed,
                                            uint32_t iter, double* out) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(h->t == 1, "single-trait handle required");
    JW_CUDA(cudaSetDevice(h->device));
    if (ensure_cap(&h->d_ve, &h->cap_ve, (size_t)h->p)) return 10;
    jw_k_bayesb_var<<<(unsigned)ceil_div(h->p, 256), 256, 0, h->stream>>>(h->d_beta, h->p, df, scale, seed, iter, h->d_ve);
    JW_LAUNCH_CHECK(h);
    if (out) JW_CUDA(cudaMemcpyAsync(out, h->d_ve, h->p * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int jwas_accumulate(jwas_handle* h, double nsamples, int bayesr) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(nsamples >= 1.0, "nsamples must be >= 1");
    JW_CUDA(cudaSetDevice(h->device));
    int64_t tp = (int64_t)h->t * h->p;
    jw_k_accumulate<<<(unsigned)ceil_div(tp, 256), 256, 0, h->stream>>>(h->d_alpha, h->d_delta, tp, nsamples, bayesr,
                                                                      h->d_mean_alpha, h->d_mean_alpha2, h->d_mean_delta);
    JW_LAUNCH_CHECK(h);
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int jwas_get_means(jwas_handle* h, float* ma, float* ma2, float* md) {
    JW_REQUIRE(h, "null handle");
    JW_CUDA(cudaSetDevice(h->device));
    size_t tp = (size_t)h->t * h->p;
    if (ma) JW_CUDA(cudaMemcpyAsync(ma, h->d_mean_alpha, tp * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (ma2) JW_CUDA(cudaMemcpyAsync(ma2, h->d_mean_alpha2, tp * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (md) JW_CUDA(cudaMemcpyAsync(md, h->d_mean_delta, tp * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int jwas_nccl_unique_id(uint8_t* out128) {
    JW_REQUIRE(out128, "jwas_nccl_unique_id: null argument");
    jw_nccl_api* N = jw_nccl();
    JW_REQUIRE(N, "libnccl.so.2 could not be loaded");
    jw_nccl_id id;
    JW_NCCL(N->GetUniqueId(&id));
    memcpy(out128, id.internal, JW_NCCL_UNIQUE_ID_BYTES);
    return 0;
}
// Turns the handle into rank `rank` of a `world`-way row-sharded sweep.  A handle created with the full matrix
// keeps only its own rows from here on (the rest of the image is freed); a handle created as a shard
// (jwas_create_shard / jwas_create_synthetic_shard with the range of jwas_shard_range) is checked against it.
// Marker statistics are (re)computed from integer counts summed over the ranks.  Call before jwas_set_blocks.
extern "C" int jwas_init_sharding(jwas_handle* h, int rank, int world, const uint8_t* unique_id128) {
    JW_REQUIRE(h, "null handle");
    JW_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "jwas_init_sharding: bad rank/world (1..8 ranks)");
    JW_REQUIRE(h->nblocks == 0, "jwas_init_sharding must be called before jwas_set_blocks");
    JW_REQUIRE(h->world == 1 && !h->nccl_comm, "jwas_init_sharding: the handle is already sharded");
    JW_CUDA(cudaSetDevice(h->device));
    h->shard_bounds.assign(world + 1, 0);
    for (int r = 0; r < world; ++r) { int64_t b, e; shard_range(h->n, r, world, &b, &e); h->shard_bounds[r] = b; h->shard_bounds[r + 1] = e; }
    const int64_t rb = h->shard_bounds[rank], re = h->shard_bounds[rank + 1];
    JW_REQUIRE(re > rb, "jwas_init_sharding: more ranks than 64-individual words");
    const bool full = h->row_begin == 0 && h->row_end == h->n;
    JW_REQUIRE(full || (h->row_begin == rb && h->row_end == re),
               "jwas_init_sharding: the handle's row range is not this rank's shard (see jwas_shard_range)");
    if (world == 1) { h->rank = 0; return 0; }
    JW_REQUIRE(unique_id128, "jwas_init_sharding: unique id required for world > 1");
    JW_REQUIRE(!h->ext_means, "jwas_init_sharding: call jwas_set_marker_means after sharding");
    jw_nccl_api* N = jw_nccl();
    JW_REQUIRE(N, "libnccl.so.2 could not be loaded");
    jw_nccl_id id;
    memcpy(id.internal, unique_id128, JW_NCCL_UNIQUE_ID_BYTES);
    JW_NCCL(N->CommInitRank(&h->nccl_comm, world, id, rank));
    h->rank = rank; h->world = world;
    if (full) {
        // keep rows [rb, re) only
        const int64_t nloc = re - rb;
        const int64_t new_pitch = ceil_div((nloc + 3) / 4, 16) * 16;
        uint8_t* d_new = nullptr;
        JW_CUDA(cudaMalloc((void**)&d_new, (size_t)h->p * new_pitch));
        JW_CUDA(cudaMemsetAsync(d_new, 0, (size_t)h->p * new_pitch, h->stream));
        JW_CUDA(cudaMemcpy2DAsync(d_new, new_pitch, h->d_packed + (rb >> 2), h->stride_d, (nloc + 3) / 4, h->p,
                                  cudaMemcpyDeviceToDevice, h->stream));
        JW_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_packed);
        h->d_packed = d_new; h->stride_d = new_pitch; h->stride = (nloc + 3) / 4;
        h->row_begin = rb; h->row_end = re;
        if (marker_counts(h)) return 11;
    }
    return marker_finalize(h);
}

// ---- fused multi-GPU exchange buffers over CUDA IPC ------------------------------------------
// One allocation per rank: [ring JW_X_RING][source rank 8][slot], a slot = 2*t*maxb + t values, each value one 16-byte
// word {lo32, tag, hi32, tag} written with a single vector store over NVLink and valid when both tags match
// (8-byte halves are self-validating: no flag, no fence, no second round trip).
extern "C" int jwas_ipc_export(jwas_handle* h, uint8_t* out64) {
    JW_REQUIRE(h && out64, "jwas_ipc_export: null argument");
    JW_REQUIRE(h->nblocks > 0, "jwas_ipc_export: call jwas_set_blocks first");
    JW_CUDA(cudaSetDevice(h->device));
    if (!h->d_xbuf) {
        h->x_slot_b = (int)h->maxb;
        h->x_slot_words = (int64_t)2 * h->t * h->x_slot_b + h->t;
        h->xbuf_bytes = (size_t)JW_X_RING * 8 * h->x_slot_words * 16;
        JW_CUDA(cudaMalloc((void**)&h->d_xbuf, h->xbuf_bytes));
        JW_CUDA(cudaMemset(h->d_xbuf, 0, h->xbuf_bytes));
    }
    cudaIpcMemHandle_t hd;
    JW_CUDA(cudaIpcGetMemHandle(&hd, h->d_xbuf));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out64, &hd, 64);
    return 0;
}
extern "C" int jwas_ipc_import(jwas_handle* h, const uint8_t* handles /* world * 64 bytes, rank order */) {
    JW_REQUIRE(h && handles, "jwas_ipc_import: null argument");
    JW_REQUIRE(h->world > 1 && h->world <= 8, "jwas_ipc_import: call jwas_init_sharding first (2..8 ranks)");
    JW_REQUIRE(h->d_xbuf, "jwas_ipc_import: call jwas_ipc_export first");
    JW_CUDA(cudaSetDevice(h->device));
    long long* slots[8];
    for (int r = 0; r < h->world; ++r) {
        void* base = h->d_xbuf;
        if (r != h->rank) {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, handles + (size_t)r * 64, 64);
            JW_CUDA(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
            h->peer_bufs[r] = base;
        }
        slots[r] = (long long*)base;
    }
    if (!h->d_peer_slots) JW_CUDA(cudaMalloc((void**)&h->d_peer_slots, 8 * sizeof(long long*)));
    JW_CUDA(cudaMemcpy(h->d_peer_slots, slots, h->world * sizeof(long long*), cudaMemcpyHostToDevice));
    h->ipc_ready = 1;
    return 0;
}
extern "C" int jwas_get_row_range(jwas_handle* h, int64_t* begin, int64_t* end) {
    JW_REQUIRE(h && begin && end, "null argument");
    *begin = h->row_begin; *end = h->row_end;
    return 0;
}

extern "C" int64_t jwas_kernel_launches(jwas_handle* h) { return h ? h->launches : 0; }
extern "C" double jwas_last_sweep_ms(jwas_handle* h) { return h ? h->last_sweep_ms : 0.0; }
extern "C" double jwas_last_stream_kernel_ms(jwas_handle* h, int64_t* launches) {
    if (!h) return 0.0;
    if (launches) *launches = h->prof_launches;
    return h->prof_ms;
}
extern "C" int jwas_get_phase_ns(jwas_handle* h, uint64_t* out24) {
    JW_REQUIRE(h && out24, "jwas_get_phase_ns: null argument");
    JW_CUDA(cudaSetDevice(h->device));
    JW_CUDA(cudaMemcpyAsync(out24, h->d_counters + 32, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    JW_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" void* jwas_stream(jwas_handle* h) { return h ? (void*)h->stream : nullptr; }
extern "C" int jwas_set_option(jwas_handle* h, const char* key, int64_t value) {
    JW_REQUIRE(h && key, "jwas_set_option: null argument");
    if (!strcmp(key, "profile")) { h->opt_profile = value; return 0; }
    if (!strcmp(key, "timers")) { h->opt_timers = value; return 0; }
    if (!strcmp(key, "gram_popcount")) { h->opt_gram_popc = value; return 0; }   // 1 = popcount kernel instead of the GEMM
    if (!strcmp(key, "lag")) {
        JW_REQUIRE(value >= 0 && value <= JW_MAX_LAG, "lag must be 0, 1 or 2");
        JW_CUDA(cudaSetDevice(h->device));
        h->opt_lag = value;
        if (h->nblocks > 0 && h->gramx_built < value) return build_all_gram(h, (int)value);    // cross-Gram on demand
        return 0;
    }
    if (!strcmp(key, "gather")) { h->opt_gather = value != 0; return 0; }
    if (!strcmp(key, "stream_variant")) { JW_REQUIRE(value >= 0 && value <= 3, "stream_variant must be 0..3"); h->opt_stream_variant = value; return 0; }
    if (!strcmp(key, "poll_ns_stream")) { JW_REQUIRE(value >= 0 && value <= 100000, "poll_ns_stream out of range"); h->opt_poll_ns_stream = value; return 0; }
    if (!strcmp(key, "poll_ns_chain")) { JW_REQUIRE(value >= 0 && value <= 100000, "poll_ns_chain out of range"); h->opt_poll_ns_chain = value; return 0; }
    if (!strcmp(key, "ws")) { h->opt_ws = value != 0; return 0; }
    if (!strcmp(key, "l2_prefetch")) { h->opt_l2_prefetch = value != 0; return 0; }
    if (!strcmp(key, "stream_pf")) { JW_REQUIRE(value >= 0 && value <= 64, "stream_pf must be 0..64"); h->opt_stream_pf = value; return 0; }
    if (!strcmp(key, "chain_ctas")) {
        JW_REQUIRE(value >= 0 && value <= 16, "chain_ctas must be in 0..16");
        JW_CUDA(cudaSetDevice(h->device));
        const bool changed = h->opt_chain_ctas != value;
        h->opt_chain_ctas = value;
        if (changed && h->nblocks > 0) return jw_fused_prepare(h);                      // slices are re-cut
        return 0;
    }
    if (!strcmp(key, "engine")) { JW_REQUIRE(value == 0 || value == 1, "engine must be 0 or 1"); h->opt_engine = value; return 0; }
    jw_set_error(std::string("unknown option: ") + key);
    return 2;
}
