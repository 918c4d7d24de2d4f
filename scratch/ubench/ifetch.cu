// Is straight-line code executed by ONE warp instruction-fetch-bound on B200?  Block A (N fully unrolled DFMAs on 4 independent
// chains) alternates with block B (same size, evicts A from the small per-scheduler instruction cache); cycles of A per pass
// against the issue-bound ideal (2.1 cycles per DFMA).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch ifetch.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N, bool WITH_B>
__global__ void k(double* out, double b, double c, double b2, long long* cyc) {
  double a0 = threadIdx.x, a1 = 1 + threadIdx.x, a2 = 2 + threadIdx.x, a3 = 3 + threadIdx.x;
  for (int pass = 0; pass < 6; ++pass) {
    const long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N / 4; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[pass] = t1 - t0;
    if (WITH_B) {
#pragma unroll
      for (int i = 0; i < N / 4; ++i) { a0 = fma(a0, b2, c); a1 = fma(a1, b2, b); a2 = fma(a2, b2, c); a3 = fma(a3, b2, b); }
    }
  }
  out[threadIdx.x] = a0 + a1 + a2 + a3;
}

template <int N, bool WITH_B> void run(const char* name) {
  double* out; long long* cyc, h[6];
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  k<N, WITH_B><<<1, 32>>>(out, 1.0000001, 1e-9, 0.9999999, cyc);
  cudaMemcpy(h, cyc, 48, cudaMemcpyDeviceToHost);
  printf("%s N=%d: cycles per DFMA by pass:", name, N);
  for (int p = 0; p < 6; ++p) printf(" %.2f", (double)h[p] / N);
  printf("\n");
}

int main() {
  run<256, false>("A only      ");
  run<1024, false>("A only      ");
  run<4096, false>("A only      ");
  run<256, true>("A then B    ");
  run<1024, true>("A then B    ");
  run<2048, true>("A then B    ");
  run<4096, true>("A then B    ");
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
