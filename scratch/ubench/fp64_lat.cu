// Latency / issue cost of the fp64 instructions on the Cholesky pivot chain (B200): dependent DFMA / DMUL chains, 1 .. 8
// independent chains per warp, F2F + MUFU.RSQ + F2F round trip, fp64 DMMA m8n8k4 dependent chain.  One warp, clock64().
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH>
__global__ void dfma_chain(double* out, double b, double c, int n, long long* cyc) {
  double a[NCH];
#pragma unroll
  for (int q = 0; q < NCH; ++q) a[q] = threadIdx.x + q;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int q = 0; q < NCH; ++q) a[q] = fma(a[q], b, c);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int q = 0; q < NCH; ++q) s += a[q];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rsqrt_chain(double* out, int n, long long* cyc) {
  double d = 2.0 + threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) d = (double)rsqrtf((float)d) + 1.5;
  const long long t1 = clock64();
  out[threadIdx.x] = d;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void dmma_chain(double* out, int n, long long* cyc) {
  double c0 = 0, c1 = 0, a = 1e-3 * threadIdx.x, b = 1e-3;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  const long long t1 = clock64();
  out[threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc, h;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  const int n = 4096;
#define RUN(K, NCH) { K<<<1, 32>>>(out, 1.0000001, 1e-9, n, cyc); K<<<1, 32>>>(out, 1.0000001, 1e-9, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("DFMA %d independent chain(s): %.1f cycles per step, %.1f per instruction\n", NCH, (double)h / n, (double)h / n / NCH); }
  RUN(dfma_chain<1>, 1) RUN(dfma_chain<2>, 2) RUN(dfma_chain<4>, 4) RUN(dfma_chain<8>, 8) RUN(dfma_chain<16>, 16)
  rsqrt_chain<<<1, 32>>>(out, n, cyc); rsqrt_chain<<<1, 32>>>(out, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("F2F.F32.F64 -> MUFU.RSQ -> F2F.F64.F32 -> DADD: %.1f cycles per round\n", (double)h / n);
  dmma_chain<<<1, 32>>>(out, n, cyc); dmma_chain<<<1, 32>>>(out, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA.8x8x4 dependent chain: %.1f cycles per instruction\n", (double)h / n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
