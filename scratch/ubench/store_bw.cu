// Per-SM global-store throughput on B200: st.global.v4 from registers vs cp.async.bulk (TMA 1-D) from shared memory,
// with few or all SMs active.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) stg_kernel(float4* out, size_t per_cta_f4, int reps, long long* cycles) {
  float4* dst = out + (size_t)blockIdx.x * per_cta_f4;
  const float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r)
    for (size_t i = threadIdx.x; i < per_cta_f4; i += 256) dst[i] = v;
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread issues bulk stores of `chunk` bytes from a 64 KB shared buffer; at most `depth` groups in flight (read side)
__global__ void __launch_bounds__(256, 1) tma_kernel(char* out, size_t per_cta_bytes, int chunk, int reps, long long* cycles) {
  extern __shared__ __align__(128) char buf[];
  for (int i = threadIdx.x; i < 65536 / 4; i += 256) reinterpret_cast<float*>(buf)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  char* dst = out + (size_t)blockIdx.x * per_cta_bytes;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int r = 0; r < reps; ++r)
      for (size_t off = 0; off < per_cta_bytes; off += chunk) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_u32(buf + (off & 65535))), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
      }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t per_cta = 1 << 20;   // 1 MB per CTA per rep
  const int reps = 8;
  char* out;
  long long* cyc;
  cudaMalloc(&out, per_cta * 148);
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  long long h[148];
  for (int ctas : {1, 8, 37, 74, 148}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) stg_kernel<<<ctas, 256>>>(reinterpret_cast<float4*>(out), per_cta / 16, reps, cyc);
        else tma_kernel<<<ctas, 256, 65536>>>(out, per_cta, mode == 1 ? 4096 : 16384, reps, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%3d CTAs  %-22s  %.1f B/clk/SM  (%s)\n", ctas, mode == 0 ? "st.global.v4" : mode == 1 ? "cp.async.bulk 4 KB" : "cp.async.bulk 16 KB",
             (double)per_cta * reps / mx, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
