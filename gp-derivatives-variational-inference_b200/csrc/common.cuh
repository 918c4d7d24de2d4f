// Shared device/host helpers for the DSVGP hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DSVGP_OK 0
#define DSVGP_ERR_ARG -1      // bad argument (null pointer, negative size, unsupported p)
#define DSVGP_ERR_LAUNCH -2   // cudaGetLastError() after a launch
#define DSVGP_ERR_WORKSPACE -3

#define DSVGP_MAXP_FAST 3     // compile-time direction counts of the blocked assembly kernels
#define DSVGP_MAXP 64         // runtime-p fallback limit

// every kernel launch of the library goes through CHECK_LAUNCH, which also counts it (dsvgp_launch_count)
namespace dsvgp { extern unsigned long long g_launch_count; }
#define CHECK_LAUNCH()                                      \
  do {                                                      \
    ++::dsvgp::g_launch_count;                              \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return DSVGP_ERR_LAUNCH;        \
  } while (0)

namespace dsvgp {

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole CTA; result valid in thread 0 (and broadcast through smem to all when bcast).
template <typename T, int NT>
__device__ __forceinline__ T block_sum(T v, T* scratch /* >= NT/32 */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  T r = (threadIdx.x < NT / 32) ? scratch[threadIdx.x] : T(0);
  if (w == 0) r = warp_sum(r);
  return r;
}

template <typename T> __device__ __forceinline__ T dexp(T x);
template <> __device__ __forceinline__ float dexp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double dexp<double>(double x) { return exp(x); }

template <typename T> __device__ __forceinline__ T dsqrt(T x);
template <> __device__ __forceinline__ float dsqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double dsqrt<double>(double x) { return sqrt(x); }

// softplus(x) = log(1 + exp(x)), the gpytorch Positive() transform; evaluated in double.
__host__ __device__ inline double softplus_d(double x) { return x > 30.0 ? x : log1p(exp(x)); }
__host__ __device__ inline double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

}  // namespace dsvgp
