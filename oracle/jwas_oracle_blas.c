/*
 * jwas_oracle_blas.c -- dense Float32 level-1 kernels standing in for the OpenBLAS
 * sdot/saxpy the reference reaches through LinearAlgebra.dot / BLAS.axpy!
 * (BayesABC.jl:48,51,76).  TEST INFRASTRUCTURE ONLY (see jwas_oracle.h).
 * Built -O3 -march=x86-64-v3 -fopenmp (portable to the GPU box host): this is the timed CPU baseline, so it gets the
 * same treatment a BLAS would (SIMD lanes, threads over n).
 */
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int jwo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

float jwo_sdot(const float* x, const float* y, int64_t n, int nthreads) {
    float acc = 0.0f;
    if (nthreads == 1 || n < 16384) {
#pragma omp simd reduction(+:acc)
        for (int64_t i = 0; i < n; ++i) acc += x[i] * y[i];
        return acc;
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for simd reduction(+:acc) num_threads(nthreads) schedule(static)
#endif
    for (int64_t i = 0; i < n; ++i) acc += x[i] * y[i];
    return acc;
}

void jwo_saxpy(float a, const float* x, float* y, int64_t n, int nthreads) {
    if (nthreads == 1 || n < 16384) {
#pragma omp simd
        for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
        return;
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for simd num_threads(nthreads) schedule(static)
#endif
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}
