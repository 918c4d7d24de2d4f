// Fused multi-tensor Adam: the optimiser step that follows every ELBO forward+backward in the reference's training
// loop (directional_vi.py:186-199 builds two torch.optim.Adam instances over the variational parameters and the
// hyper-parameters + likelihood noise; :251-254 steps them and their per-minibatch LR schedulers).
//
// torch's default (foreach) Adam launches ~10 elementwise kernels per dtype group; at the reference's minibatch sizes
// that is comparable to the fused ELBO step itself.  Here ONE launch updates every tensor of an optimiser:
// the tensor table (<= ADAM_MAX_TENSORS entries) travels in the kernel parameters (no device-side table to keep
// alive, gradients may live at a new address every step), a block finds its tensor by a prefix-sum search, and each
// element is read and written exactly once:  param, grad, exp_avg, exp_avg_sq in; param, exp_avg, exp_avg_sq out.
// Arithmetic follows torch.optim.Adam (amsgrad=False, maximize=False, L2 weight decay) operation by operation, in
// the parameter dtype:
//   g   = grad + wd * p
//   m  += (g - m) * (1 - b1)                     (lerp_)
//   v   = v * b2 + (1 - b2) * g * g              (mul_ + addcmul_)
//   p  -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// `tri_n` > 0 marks an n x n matrix of which only the lower triangle can have a non-zero gradient
// (chol_variational_covar): rows are cut at the diagonal, which skips half the traffic and gives bit-identical results
// as long as the strictly-upper moments are zero (they are never touched by this optimiser).
#include "common.cuh"
#include "optim.cuh"

namespace dsvgp {

constexpr int ADAM_THREADS = 256, ADAM_ILP = 4, ADAM_CHUNK = ADAM_THREADS * ADAM_ILP * 4;   // elements per block

template <typename T>
struct AdamTable {
  T* p[ADAM_MAX_TENSORS];
  const T* g[ADAM_MAX_TENSORS];
  T* m[ADAM_MAX_TENSORS];
  T* v[ADAM_MAX_TENSORS];
  int64_t numel[ADAM_MAX_TENSORS];
  int tri_n[ADAM_MAX_TENSORS];
  int group[ADAM_MAX_TENSORS];
  int first_block[ADAM_MAX_TENSORS + 1];
  T lr_eff[ADAM_MAX_GROUPS], b1[ADAM_MAX_GROUPS], b2[ADAM_MAX_GROUPS], eps[ADAM_MAX_GROUPS], wd[ADAM_MAX_GROUPS],
      rsq_bc2[ADAM_MAX_GROUPS];   // lr / (1 - b1^t) ;  sqrt(1 - b2^t)
  int ntensors;
};

template <typename T> __device__ __forceinline__ T tsqrt(T x);
template <> __device__ __forceinline__ float tsqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double tsqrt<double>(double x) { return sqrt(x); }

template <typename T>
__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const __grid_constant__ AdamTable<T> tb) {
  int t = 0;
#pragma unroll 1
  while (t + 1 < tb.ntensors && (int)blockIdx.x >= tb.first_block[t + 1]) ++t;
  const int64_t base = (int64_t)(blockIdx.x - tb.first_block[t]) * ADAM_CHUNK;
  const int64_t numel = tb.numel[t];
  const int gi = tb.group[t], tri = tb.tri_n[t];
  const T lr_eff = tb.lr_eff[gi], b1 = tb.b1[gi], b2 = tb.b2[gi], eps = tb.eps[gi], wd = tb.wd[gi], sbc2 = tb.rsq_bc2[gi];
  T* __restrict__ P = tb.p[t];
  const T* __restrict__ G = tb.g[t];
  T* __restrict__ Mo = tb.m[t];
  T* __restrict__ V = tb.v[t];
#pragma unroll
  for (int q = 0; q < ADAM_ILP * 4; ++q) {
    const int64_t e = base + (int64_t)q * ADAM_THREADS + threadIdx.x;
    if (e >= numel) break;
    if (tri > 0) {
      const int r = (int)(e / tri), c = (int)(e - (int64_t)r * tri);
      if (c > r) continue;
    }
    T p = P[e], g = G[e], m = Mo[e], v = V[e];
    if (wd != T(0)) g = g + wd * p;
    m = m + (g - m) * (T(1) - b1);
    v = v * b2 + (T(1) - b2) * g * g;
    const T denom = tsqrt(v) / sbc2 + eps;
    p = p - lr_eff * (m / denom);
    P[e] = p;
    Mo[e] = m;
    V[e] = v;
  }
}

template <typename T>
int adam_step(int ntensors, const int64_t* desc_host, int ngroups, const double* group_host, cudaStream_t st) {
  if (ntensors <= 0) return DSVGP_OK;
  if (!desc_host || !group_host || ngroups <= 0 || ngroups > ADAM_MAX_GROUPS) return DSVGP_ERR_ARG;
  for (int t0 = 0; t0 < ntensors; t0 += ADAM_MAX_TENSORS) {
    AdamTable<T> tb;
    const int nt = ntensors - t0 < ADAM_MAX_TENSORS ? ntensors - t0 : ADAM_MAX_TENSORS;
    int blocks = 0;
    for (int i = 0; i < nt; ++i) {
      const int64_t* d = desc_host + (int64_t)(t0 + i) * 8;
      tb.p[i] = reinterpret_cast<T*>(d[0]);
      tb.g[i] = reinterpret_cast<const T*>(d[1]);
      tb.m[i] = reinterpret_cast<T*>(d[2]);
      tb.v[i] = reinterpret_cast<T*>(d[3]);
      tb.numel[i] = d[4];
      tb.group[i] = (int)d[5];
      tb.tri_n[i] = (int)d[6];
      if (!tb.p[i] || !tb.g[i] || !tb.m[i] || !tb.v[i] || d[4] < 0 || d[5] < 0 || d[5] >= ngroups) return DSVGP_ERR_ARG;
      if (d[6] > 0 && d[6] * d[6] != d[4]) return DSVGP_ERR_ARG;
      tb.first_block[i] = blocks;
      blocks += (int)ceil_div64(d[4], ADAM_CHUNK);
    }
    tb.first_block[nt] = blocks;
    tb.ntensors = nt;
    for (int gidx = 0; gidx < ngroups; ++gidx) {
      const double* h = group_host + gidx * 8;     // lr, beta1, beta2, eps, weight_decay, step, -, -
      const double lr = h[0], b1 = h[1], b2 = h[2], step = h[5];
      const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
      tb.lr_eff[gidx] = (T)(lr / bc1);
      tb.b1[gidx] = (T)b1;
      tb.b2[gidx] = (T)b2;
      tb.eps[gidx] = (T)h[3];
      tb.wd[gidx] = (T)h[4];
      tb.rsq_bc2[gidx] = (T)sqrt(bc2);
    }
    if (blocks == 0) continue;
    adam_kernel<T><<<blocks, ADAM_THREADS, 0, st>>>(tb);
    CHECK_LAUNCH();
  }
  return DSVGP_OK;
}

template int adam_step<float>(int, const int64_t*, int, const double*, cudaStream_t);
template int adam_step<double>(int, const int64_t*, int, const double*, cudaStream_t);

}  // namespace dsvgp
