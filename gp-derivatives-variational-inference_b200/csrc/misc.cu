// Small fused elementwise / reduction kernels around the whitening products:
// positivity transforms, predictive mean / diagonal variance, Gaussian expected log-likelihood, KL, and the
// pieces of their backward that are not GEMMs.  All HBM-bound; warp-shuffle + shared-memory reductions, partial
// sums written per CTA and summed by a second tiny kernel (deterministic, no float atomics).
//
// Reference semantics (file:line are under /root/reference/directionalvi unless noted; [GPT] = gpytorch 1.4.0):
//   mean     = A^T m + c on every output incl. derivative outputs   DirectionalGradVariationalStrategy.py:126,188
//   variance = diag(K_xx) + 1e-4 + diag(A^T (S - I) A)              :192-205, via MatmulLazyTensor.diag [GPT]
//   likelihood(dist) adds sigma^2, sigma^2 = softplus(raw)+1e-4     [GPT] GaussianLikelihood
//   expected_log_prob = -0.5*(((y-mu)^2 + var)/sigma^2 + log sigma^2 + log 2pi)   [GPT]
//   KL(q||N(0,I)) = 0.5*(|tril(Ls)|_F^2 + m.m - M' - sum log Ls_ii^2)             [GPT]
#include <cuda_fp16.h>

#include "common.cuh"
#include "misc.cuh"

namespace dsvgp {

// hyp = {ell, os, noise, c, sigmoid(raw_ell), sigmoid(raw_os), sigmoid(raw_noise), 0}
template <typename T>
__global__ void hyp_from_raw_kernel(const T* raw_ell, const T* raw_os, const T* raw_noise, const T* c, double* hyp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double re = (double)raw_ell[0];
  hyp[0] = softplus_d(re);
  hyp[4] = sigmoid_d(re);
  const double ro = raw_os ? (double)raw_os[0] : 0.0;
  hyp[1] = raw_os ? softplus_d(ro) : 1.0;
  hyp[5] = raw_os ? sigmoid_d(ro) : 0.0;
  const double rn = raw_noise ? (double)raw_noise[0] : 0.0;
  hyp[2] = raw_noise ? softplus_d(rn) + 1e-4 : 0.0;
  hyp[6] = raw_noise ? sigmoid_d(rn) : 0.0;
  hyp[3] = c ? (double)c[0] : 0.0;
  hyp[7] = 0.0;
}

__global__ void pad_identity_kernel(double* A, int64_t ld, int Mq, int Mp) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int i = (int)(e / Mp), j = (int)(e % Mp);
  if (i >= Mq || j >= Mq) A[(int64_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
}

template <typename S, typename D>
__global__ void cast2d_kernel(const S* __restrict__ src, int64_t lds, D* __restrict__ dst, int64_t ldd, int rows,
                              int cols, int tril) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= cols || i >= rows) return;
  dst[(int64_t)i * ldd + j] = (tril && j > i) ? D(0) : (D)src[(int64_t)i * lds + j];
}

// fill the strictly-upper triangle from the lower one
template <typename T>
__global__ void mirror_lower_kernel(T* A, int64_t ld, int n) {
  __shared__ T tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < n && j < n) ? A[(int64_t)i * ld + j] : T(0);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bj * 32 + r, j = bi * 32 + tx;      // transposed block position
    if (i < n && j < n && j > i) A[(int64_t)i * ld + j] = tile[tx][r];
  }
}

template <typename T>
__global__ void add_outer_kernel(T* A, int64_t ld, int n, const T* __restrict__ u, const T* __restrict__ v, T alpha) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n || i >= n) return;
  A[(int64_t)i * ld + j] += alpha * u[i] * v[j];
}

template <typename T>
__global__ void tril_minus_eye_kernel(const T* __restrict__ Ls, int64_t ldl, T* __restrict__ E, int64_t lde, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n || i >= n) return;
  E[(int64_t)i * lde + j] = (j > i) ? T(0) : Ls[(int64_t)i * ldl + j] - (i == j ? T(1) : T(0));
}

// Psi = 0.5*(Phi + Phi^T), Phi = tril(Y) with halved diagonal  (symmetrised Cholesky-backward middle factor)
__global__ void sym_phi_kernel(const double* __restrict__ Y, int64_t ldy, double* __restrict__ P, int64_t ldp, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n || i >= n) return;
  const int a = max(i, j), b = min(i, j);
  P[(int64_t)i * ldp + j] = 0.5 * Y[(int64_t)a * ldy + b];
}

// Phi = tril(Y) with halved diagonal (Cholesky-backward middle factor, Murray 2016)
__global__ void phi_lower_kernel(const double* __restrict__ Y, int64_t ldy, double* __restrict__ P, int64_t ldp, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n || i >= n) return;
  P[(int64_t)i * ldp + j] = (j < i) ? Y[(int64_t)i * ldy + j] : (j == i ? 0.5 * Y[(int64_t)i * ldy + j] : 0.0);
}

// Lower triangle of Phi(X + u v^T) in fp64 from X in the model dtype, two columns per thread: the add_outer -> cast2d -> phi_lower
// sequence of the Cholesky-backward tail as ONE pass that reads and writes the lower triangle only (the products that consume P
// take it with a lower-triangle flag and never load the other side): 3 launches / ~260 MB of traffic -> 1 launch / ~60 MB at M' = 3072.
template <typename T>
__global__ void phi_outer_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ u, const T* __restrict__ v,
                                 double* __restrict__ P, int64_t ldp, int n) {
  const int i = blockIdx.y, j = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n || j > i) return;
  const T ui = u[i];
  const T x0 = X[(int64_t)i * ldx + j] + T(1) * ui * v[j];
  double p0 = (double)x0, p1 = 0.0;
  if (j == i) p0 *= 0.5;
  if (j + 1 <= i) {
    const T x1 = X[(int64_t)i * ldx + j + 1] + T(1) * ui * v[j + 1];
    p1 = (j + 1 == i) ? 0.5 * (double)x1 : (double)x1;
  }
  double* dst = P + (int64_t)i * ldp + j;
  if (((ldp & 1) == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0) && j + 1 < n) {
    *reinterpret_cast<double2*>(dst) = make_double2(p0, p1);      // (the entry right of the diagonal is written as 0)
  } else {
    dst[0] = p0;
    if (j + 1 < n) dst[1] = p1;
  }
}

// A <- (A + A^T) / 2 in place
__global__ void symmetrize_kernel(double* A, int64_t ld, int n) {
  __shared__ double ta[32][33], tb[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    ta[r][tx] = (i < n && j < n) ? A[(int64_t)i * ld + j] : 0.0;
    const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
    tb[r][tx] = (i2 < n && j2 < n) ? A[(int64_t)i2 * ld + j2] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < n && j < n) A[(int64_t)i * ld + j] = 0.5 * (ta[r][tx] + tb[tx][r]);
    const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
    if (bi != bj && i2 < n && j2 < n) A[(int64_t)i2 * ld + j2] = 0.5 * (tb[r][tx] + ta[tx][r]);
  }
}

// --------------------------------------------------------------------------------- predictive mean / variance
// partial column sums over a slab of rows:  pm[s][j] = sum_i A_ij m_i ,  pv[s][j] = sum_i A_ij C_ij
template <typename T>
__global__ void __launch_bounds__(256)
col_dots_kernel(const T* __restrict__ A, const T* __restrict__ C, int64_t ld, int rows, int nq,
                const T* __restrict__ m, int rows_per_slab, T* __restrict__ pm, T* __restrict__ pv,
                unsigned* __restrict__ cmax_bits) {
  __shared__ T sm[4][64], sv[4][64];
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + cl;
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(rows, r0 + rows_per_slab);
  T am = 0, av = 0;
  float cm = 0.f;                       // max|C| seen by this thread (scale of the dA operand, csrc/tc_prep.cu)
  if (j < nq) {
    int i = r0 + rl;
    for (; i + 12 < r1; i += 16) {
      T a[4], c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = A[(int64_t)(i + 4 * q) * ld + j];
        c[q] = C ? C[(int64_t)(i + 4 * q) * ld + j] : T(0);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        am += a[q] * m[i + 4 * q];
        av += a[q] * c[q];
        cm = fmaxf(cm, fabsf((float)c[q]));
      }
    }
    for (; i < r1; i += 4) {
      const T a = A[(int64_t)i * ld + j];
      am += a * m[i];
      if (C) {
        const T c = C[(int64_t)i * ld + j];
        av += a * c;
        cm = fmaxf(cm, fabsf((float)c));
      }
    }
  }
  if (cmax_bits != nullptr) {           // order-independent maximum on the bit pattern: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if ((threadIdx.x & 31) == 0 && cm > 0.f) atomicMax(cmax_bits, __float_as_uint(cm));
  }
  sm[rl][cl] = am;
  sv[rl][cl] = av;
  __syncthreads();
  if (rl == 0 && j < nq) {
    pm[(int64_t)blockIdx.y * nq + j] = sm[0][cl] + sm[1][cl] + sm[2][cl] + sm[3][cl];
    pv[(int64_t)blockIdx.y * nq + j] = sv[0][cl] + sv[1][cl] + sv[2][cl] + sv[3][cl];
  }
}

// 3xFP16 training step: A exists only as the two-half split (hi, lo) of A * *a_scale written by the whitening product (its fp32
// copy is not written at all: half of that product's store phase).  Same partial sums as col_dots_kernel with
// A_ij = (hi_ij + lo_ij) / a_scale; a thread owns two adjacent columns (half2 / float2 accesses).
__global__ void __launch_bounds__(256)
col_dots_h_kernel(const __half* __restrict__ Ah, const __half* __restrict__ Al, int64_t ldh, const float* __restrict__ a_scale,
                  const float* __restrict__ C, int64_t ld, int rows, int nq, const float* __restrict__ m, int rows_per_slab,
                  float* __restrict__ pm, float* __restrict__ pv, unsigned* __restrict__ cmax_bits) {
  // 64 columns per CTA like col_dots_kernel (same grid, same slabs): a warp covers them as 32 x 2, the 8 warps take rows i, i + 8, ...
  __shared__ float sm[8][64], sv[8][64];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int j = blockIdx.x * 64 + 2 * cl;
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(rows, r0 + rows_per_slab);
  const float inv = 1.f / *a_scale;
  float am0 = 0.f, am1 = 0.f, av0 = 0.f, av1 = 0.f, cm = 0.f;
  if (j + 1 < nq) {
    int i = r0 + rl;
    for (; i + 24 < r1; i += 32) {
      __half2 h[4], l[4];
      float2 c[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[q] = *reinterpret_cast<const __half2*>(Ah + (int64_t)(i + 8 * q) * ldh + j);
        l[q] = *reinterpret_cast<const __half2*>(Al + (int64_t)(i + 8 * q) * ldh + j);
        c[q] = *reinterpret_cast<const float2*>(C + (int64_t)(i + 8 * q) * ld + j);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 hf = __half22float2(h[q]), lf = __half22float2(l[q]);
        const float a0 = (hf.x + lf.x) * inv, a1 = (hf.y + lf.y) * inv, mi = m[i + 8 * q];
        am0 += a0 * mi; am1 += a1 * mi;
        av0 += a0 * c[q].x; av1 += a1 * c[q].y;
        cm = fmaxf(cm, fmaxf(fabsf(c[q].x), fabsf(c[q].y)));
      }
    }
    for (; i < r1; i += 8) {
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(Ah + (int64_t)i * ldh + j));
      const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(Al + (int64_t)i * ldh + j));
      const float2 c = *reinterpret_cast<const float2*>(C + (int64_t)i * ld + j);
      const float a0 = (hf.x + lf.x) * inv, a1 = (hf.y + lf.y) * inv, mi = m[i];
      am0 += a0 * mi; am1 += a1 * mi;
      av0 += a0 * c.x; av1 += a1 * c.y;
      cm = fmaxf(cm, fmaxf(fabsf(c.x), fabsf(c.y)));
    }
  } else if (j < nq) {                  // odd last column
    for (int i = r0 + rl; i < r1; i += 8) {
      const float a0 = (__half2float(Ah[(int64_t)i * ldh + j]) + __half2float(Al[(int64_t)i * ldh + j])) * inv;
      const float c = C[(int64_t)i * ld + j];
      am0 += a0 * m[i];
      av0 += a0 * c;
      cm = fmaxf(cm, fabsf(c));
    }
  }
  if (cmax_bits != nullptr) {           // order-independent maximum on the bit pattern: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if ((threadIdx.x & 31) == 0 && cm > 0.f) atomicMax(cmax_bits, __float_as_uint(cm));
  }
  sm[rl][2 * cl] = am0; sm[rl][2 * cl + 1] = am1;
  sv[rl][2 * cl] = av0; sv[rl][2 * cl + 1] = av1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c_ = threadIdx.x, jc = blockIdx.x * 64 + c_;
    if (jc < nq) {
      float a = 0.f, v = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) { a += sm[r][c_]; v += sv[r][c_]; }
      pm[(int64_t)blockIdx.y * nq + jc] = a;
      pv[(int64_t)blockIdx.y * nq + jc] = v;
    }
  }
}

// partial column sums  pv[s][j] = sum_i B'_ij (2 A_ij + B'_ij) = sum_i (B_ij^2 - A_ij^2) with B' = B - A
// (prediction path: no backward, so C is never formed)
template <typename T>
__global__ void __launch_bounds__(256)
col_sqdiff_kernel(const T* __restrict__ A, const T* __restrict__ B, int64_t ld, int rows, int nq,
                  const T* __restrict__ m, int rows_per_slab, T* __restrict__ pm, T* __restrict__ pv) {
  __shared__ T sm[4][64], sv[4][64];
  const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + cl;
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(rows, r0 + rows_per_slab);
  T am = 0, av = 0;
  if (j < nq) {
    for (int i = r0 + rl; i < r1; i += 4) {
      const T a = A[(int64_t)i * ld + j], bp = B[(int64_t)i * ld + j];    // bp = B' = (L_s^T - I) A
      am += a * m[i];
      av += bp * (a + a + bp);
    }
  }
  sm[rl][cl] = am;
  sv[rl][cl] = av;
  __syncthreads();
  if (rl == 0 && j < nq) {
    pm[(int64_t)blockIdx.y * nq + j] = sm[0][cl] + sm[1][cl] + sm[2][cl] + sm[3][cl];
    pv[(int64_t)blockIdx.y * nq + j] = sv[0][cl] + sv[1][cl] + sv[2][cl] + sv[3][cl];
  }
}

template <typename T>
__global__ void predict_finish_kernel(const T* __restrict__ pm, const T* __restrict__ pv, int nslab, int nq, int p2,
                                      const double* __restrict__ hyp, double pred_jitter, int add_noise,
                                      double min_var, T* __restrict__ mu, T* __restrict__ var) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nq) return;
  double sm = 0, sv = 0;
  for (int s = 0; s < nslab; ++s) {
    sm += (double)pm[(int64_t)s * nq + j];
    sv += (double)pv[(int64_t)s * nq + j];
  }
  const double ell = hyp[0], os = hyp[1];
  const double kd = (j % (p2 + 1)) == 0 ? os : os / (ell * ell);
  double v = kd + pred_jitter + sv + (double)add_noise * hyp[2];      // add_noise = how many times likelihood() was applied (Q3)
  if (v < min_var) v = min_var;
  mu[j] = (T)(sm + hyp[3]);
  var[j] = (T)v;
}

// ------------------------------------------------------------------------------------------ ELBO data term
// sc_part[b] = {sum_j term_j * w, explicit d/dsigma^2}, gmu_j = dELBO/dmu_j, gvar_j = dELBO/dvar_j
template <typename T>
__global__ void __launch_bounds__(256)
elbo_terms_kernel(const T* __restrict__ mu, const T* __restrict__ var, const T* __restrict__ y, int nq,
                  const double* __restrict__ hyp, double w, double min_var, T* __restrict__ gmu,
                  T* __restrict__ gvar, double* __restrict__ sc_part) {
  __shared__ double red[64];
  const int j = blockIdx.x * 256 + threadIdx.x;
  const double s2 = hyp[2];
  double term = 0, ds2 = 0;
  if (j < nq) {
    const double r = (double)y[j] - (double)mu[j], v = (double)var[j];
    const double quad = r * r + v;
    term = -0.5 * (quad / s2 + log(s2) + 1.8378770664093453) * w;
    ds2 = 0.5 * (quad / (s2 * s2) - 1.0 / s2) * w;
    gmu[j] = (T)(w * r / s2);
    gvar[j] = (T)((v > min_var) ? -0.5 * w / s2 : 0.0);
  }
  const double a = block_sum<double, 256>(term, red);
  const double b = block_sum<double, 256>(ds2, red + 32);
  if (threadIdx.x == 0) {
    sc_part[2 * blockIdx.x] = a;
    sc_part[2 * blockIdx.x + 1] = b;
  }
}

// PredictiveLogLikelihood data term (gpytorch.mlls.PredictiveLogLikelihood, mll_type="PLL", directional_vi.py:218-219):
// log N(y_j; mu_j, v_j) with v the marginal variance (the caller has already added the noise as many times as
// likelihood() was applied -- twice in the reference loop, Q3).  No explicit noise derivative: it flows through v.
template <typename T>
__global__ void __launch_bounds__(256)
pll_terms_kernel(const T* __restrict__ mu, const T* __restrict__ var, const T* __restrict__ y, int nq, double w,
                 double min_var, T* __restrict__ gmu, T* __restrict__ gvar, double* __restrict__ sc_part) {
  __shared__ double red[32];
  const int j = blockIdx.x * 256 + threadIdx.x;
  double term = 0;
  if (j < nq) {
    const double r = (double)y[j] - (double)mu[j], v = (double)var[j];
    term = -0.5 * (r * r / v + log(v) + 1.8378770664093453) * w;
    gmu[j] = (T)(w * r / v);
    gvar[j] = (T)((v > min_var) ? -0.5 * w * (1.0 / v - r * r / (v * v)) : 0.0);
  }
  const double a = block_sum<double, 256>(term, red);
  if (threadIdx.x == 0) {
    sc_part[2 * blockIdx.x] = a;
    sc_part[2 * blockIdx.x + 1] = 0.0;
  }
}

// out[k] += sum_b part[b*stride + k], k < width   (one CTA)
__global__ void sum_scalar_parts_kernel(const double* __restrict__ part, int nparts, int stride, int width,
                                        double* __restrict__ out) {
  __shared__ double red[32];
  for (int k = 0; k < width; ++k) {
    double s = 0;
    for (int b = threadIdx.x; b < nparts; b += 256) s += part[(int64_t)b * stride + k];
    s = block_sum<double, 256>(s, red);
    if (threadIdx.x == 0) out[k] += s;
    __syncthreads();
  }
}

// gsc += {d ell, d os, d sigma^2, d c} contributions that flow through mean (+c) and the K_xx diagonal
template <typename T>
__global__ void __launch_bounds__(256)
pred_bwd_scalars_kernel(const T* __restrict__ gmu, const T* __restrict__ gvar, int nq, int p2,
                        const double* __restrict__ hyp, int add_noise, double* __restrict__ part) {
  __shared__ double red[32];
  const double ell = hyp[0], os = hyp[1];
  double d_ell = 0, d_os = 0, d_s2 = 0, d_c = 0;
  for (int j = blockIdx.x * 256 + threadIdx.x; j < nq; j += gridDim.x * 256) {
    const double gv = (double)gvar[j];
    const bool val = (j % (p2 + 1)) == 0;
    d_os += gv * (val ? 1.0 : 1.0 / (ell * ell));
    d_ell += val ? 0.0 : gv * (-2.0 * os / (ell * ell * ell));
    d_s2 += (double)add_noise * gv;
    d_c += (double)gmu[j];
  }
  double v[4] = {d_ell, d_os, d_s2, d_c};
  for (int k = 0; k < 4; ++k) {
    const double s = block_sum<double, 256>(v[k], red);
    if (threadIdx.x == 0) part[4 * blockIdx.x + k] = s;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ backward through mean / variance
// In place  C_ij <- m_i*gmu_j + 2*gvar_j*C_ij  (= dELBO/dA),  Ag_ij <- A_ij*gvar_j  (left factor of the weighted
// SYRK G = A diag(gvar) A^T),  tp[s][i] = sum_{j in slab s} A_ij*gmu_j  (= dELBO/dm and the rank-one part of dL).
__device__ __forceinline__ float tf32_lo(float x) {     // rn_tf32(x - trunc_tf32(x)), see trmm_tc.cu
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(r));
  return __uint_as_float(t);
}

template <typename T>
__global__ void __launch_bounds__(256)
dA_kernel(const T* __restrict__ A, T* __restrict__ C, T* __restrict__ Ag, int64_t ld, int rows, int nq,
          const T* __restrict__ m, const T* __restrict__ gmu, const T* __restrict__ gvar, int cols_per_slab,
          T* __restrict__ tp, T* __restrict__ Clo, T* __restrict__ Aglo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + warp;
  if (i >= rows) return;
  const int c0 = blockIdx.y * cols_per_slab, c1 = min(nq, c0 + cols_per_slab);
  const T mi = m[i];
  T t = 0;
  int jstart = c0;
  if constexpr (sizeof(T) == 4) {
    // 16-byte path: a lane owns 4 consecutive columns; every array is read / written once, the four outputs
    // (2.4 GB at C3) with evict-first stores -- they are consumed once by the next GEMM and exceed L2 anyway
    const bool al = ((ld & 3) == 0) && ((c0 & 3) == 0) &&
                    (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(Ag) |
                       reinterpret_cast<uintptr_t>(Clo) | reinterpret_cast<uintptr_t>(Aglo) |
                       reinterpret_cast<uintptr_t>(gmu) | reinterpret_cast<uintptr_t>(gvar)) & 15) == 0);
    if (al) {
      const int cvec = c0 + ((c1 - c0) & ~3);
      for (int j = c0 + 4 * lane; j < cvec; j += 128) {
        const int64_t o = (int64_t)i * ld + j;
        const float4 a = __ldcs(reinterpret_cast<const float4*>(A + o));
        const float4 c = __ldcs(reinterpret_cast<const float4*>(C + o));
        const float4 gm = *reinterpret_cast<const float4*>(gmu + j);
        const float4 gv = *reinterpret_cast<const float4*>(gvar + j);
        t += a.x * gm.x + a.y * gm.y + a.z * gm.z + a.w * gm.w;
        const float4 dA = make_float4(mi * gm.x + 2.f * gv.x * c.x, mi * gm.y + 2.f * gv.y * c.y,
                                      mi * gm.z + 2.f * gv.z * c.z, mi * gm.w + 2.f * gv.w * c.w);
        const float4 ag = make_float4(a.x * gv.x, a.y * gv.y, a.z * gv.z, a.w * gv.w);
        __stcs(reinterpret_cast<float4*>(C + o), dA);
        if (Ag) __stcs(reinterpret_cast<float4*>(Ag + o), ag);
        if (Clo) __stcs(reinterpret_cast<float4*>(Clo + o), make_float4(tf32_lo(dA.x), tf32_lo(dA.y), tf32_lo(dA.z), tf32_lo(dA.w)));
        if (Aglo) __stcs(reinterpret_cast<float4*>(Aglo + o), make_float4(tf32_lo(ag.x), tf32_lo(ag.y), tf32_lo(ag.z), tf32_lo(ag.w)));
      }
      jstart = cvec;
    }
  }
  for (int j = jstart + lane; j < c1; j += 32) {
    const int64_t o = (int64_t)i * ld + j;
    const T a = A[o], gm = gmu[j], gv = gvar[j];
    t += a * gm;
    const T dA = mi * gm + T(2) * gv * C[o];
    C[o] = dA;
    if (Ag) Ag[o] = a * gv;
    if constexpr (sizeof(T) == 4) {
      if (Clo) Clo[o] = tf32_lo(dA);
      if (Aglo) Aglo[o] = tf32_lo(a * gv);
    }
  }
  t = warp_sum(t);
  if (lane == 0) tp[(int64_t)blockIdx.y * rows + i] = t;
}

// 3xFP16 path: the same pass, but dA and A_g leave only as the two-half splits of dA * *s_dA and A_g * *s_Ag (the
// operands of dK_zx = W^T dA and G = A_g A^T): 8 bytes read and 8 bytes written per element instead of 8 and 16.
__global__ void __launch_bounds__(256)
dA_half_kernel(const float* __restrict__ A, const float* __restrict__ C, int64_t ld, int rows, int nq,
               const float* __restrict__ m, const float* __restrict__ gmu, const float* __restrict__ gvar, int cols_per_slab,
               float* __restrict__ tp, __half* __restrict__ dAh, __half* __restrict__ dAl, __half* __restrict__ Agh,
               __half* __restrict__ Agl, int64_t ldh, const float* __restrict__ s_dA, const float* __restrict__ s_Ag,
               const __half* __restrict__ Ah, const __half* __restrict__ Al, const float* __restrict__ a_scale) {
  // A comes as the fp32 matrix, or (A == nullptr) as the two-half split (Ah, Al; leading dimension ldh) of A * *a_scale
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + warp;
  if (i >= rows) return;
  const int c0 = blockIdx.y * cols_per_slab, c1 = min(nq, c0 + cols_per_slab);
  const float mi = m[i], sd = *s_dA, sg = *s_Ag;
  const float ainv = A ? 1.f : 1.f / *a_scale;
  float t = 0.f;
  auto put = [&](int64_t oh, float dA, float ag) {
    const float x = dA * sd, y = ag * sg;
    const __half xh = __float2half_rn(x), yh = __float2half_rn(y);
    dAh[oh] = xh;
    dAl[oh] = __float2half_rn(x - __half2float(xh));
    Agh[oh] = yh;
    Agl[oh] = __float2half_rn(y - __half2float(yh));
  };
  const bool al = ((ld & 3) == 0) && ((ldh & 3) == 0) && ((c0 & 3) == 0) &&
                  (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(gmu) |
                     reinterpret_cast<uintptr_t>(gvar)) & 15) == 0) &&
                  (((reinterpret_cast<uintptr_t>(Ah) | reinterpret_cast<uintptr_t>(Al)) & 7) == 0) &&
                  (((reinterpret_cast<uintptr_t>(dAh) | reinterpret_cast<uintptr_t>(dAl) | reinterpret_cast<uintptr_t>(Agh) |
                     reinterpret_cast<uintptr_t>(Agl)) & 7) == 0);
  int jstart = c0;
  if (al) {
    const int cvec = c0 + ((c1 - c0) & ~3);
    for (int j = c0 + 4 * lane; j < cvec; j += 128) {
      const int64_t o = (int64_t)i * ld + j, oh = (int64_t)i * ldh + j;
      float4 a;
      if (A) a = __ldcs(reinterpret_cast<const float4*>(A + o));
      else {
        const uint2 hv = __ldcs(reinterpret_cast<const uint2*>(Ah + oh)), lv = __ldcs(reinterpret_cast<const uint2*>(Al + oh));
        const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&hv.x)), h23 = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
        const float2 l01 = __half22float2(*reinterpret_cast<const __half2*>(&lv.x)), l23 = __half22float2(*reinterpret_cast<const __half2*>(&lv.y));
        a = make_float4((h01.x + l01.x) * ainv, (h01.y + l01.y) * ainv, (h23.x + l23.x) * ainv, (h23.y + l23.y) * ainv);
      }
      const float4 c = __ldcs(reinterpret_cast<const float4*>(C + o));
      const float4 gm = *reinterpret_cast<const float4*>(gmu + j);
      const float4 gv = *reinterpret_cast<const float4*>(gvar + j);
      t += a.x * gm.x + a.y * gm.y + a.z * gm.z + a.w * gm.w;
      const float x[4] = {(mi * gm.x + 2.f * gv.x * c.x) * sd, (mi * gm.y + 2.f * gv.y * c.y) * sd,
                          (mi * gm.z + 2.f * gv.z * c.z) * sd, (mi * gm.w + 2.f * gv.w * c.w) * sd};
      const float y[4] = {a.x * gv.x * sg, a.y * gv.y * sg, a.z * gv.z * sg, a.w * gv.w * sg};
      __half xh[4], xl[4], yh[4], yl[4];
#pragma unroll
      for (int z = 0; z < 4; ++z) {
        xh[z] = __float2half_rn(x[z]);
        xl[z] = __float2half_rn(x[z] - __half2float(xh[z]));
        yh[z] = __float2half_rn(y[z]);
        yl[z] = __float2half_rn(y[z] - __half2float(yh[z]));
      }
      __stcs(reinterpret_cast<uint2*>(dAh + oh), *reinterpret_cast<const uint2*>(xh));
      __stcs(reinterpret_cast<uint2*>(dAl + oh), *reinterpret_cast<const uint2*>(xl));
      __stcs(reinterpret_cast<uint2*>(Agh + oh), *reinterpret_cast<const uint2*>(yh));
      __stcs(reinterpret_cast<uint2*>(Agl + oh), *reinterpret_cast<const uint2*>(yl));
    }
    jstart = cvec;
  }
  for (int j = jstart + lane; j < c1; j += 32) {
    const int64_t o = (int64_t)i * ld + j;
    const float a = A ? A[o] : (__half2float(Ah[(int64_t)i * ldh + j]) + __half2float(Al[(int64_t)i * ldh + j])) * ainv;
    const float gm = gmu[j], gv = gvar[j];
    t += a * gm;
    put((int64_t)i * ldh + j, mi * gm + 2.f * gv * C[o], a * gv);
  }
  t = warp_sum(t);
  if (lane == 0) tp[(int64_t)blockIdx.y * rows + i] = t;
}

template <typename T>
__global__ void sum_parts_kernel(const T* __restrict__ part, int nparts, int len, T* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0;
  for (int b = 0; b < nparts; ++b) s += (double)part[(int64_t)b * len + i];
  out[i] = (T)s;
}

// ------------------------------------------------------------------------------------------------------ KL
template <typename T>
__global__ void __launch_bounds__(256)
kl_kernel(const T* __restrict__ m, const T* __restrict__ Ls, int64_t ld, int Mq, double* __restrict__ part) {
  __shared__ double red[32];
  double s = 0;
  const int64_t tot = (int64_t)Mq * Mq;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < tot; e += (int64_t)gridDim.x * 256) {
    const int i = (int)(e / Mq), j = (int)(e % Mq);
    if (j <= i) {
      const double v = (double)Ls[(int64_t)i * ld + j];
      s += 0.5 * v * v;
      if (i == j) s += 0.5 * ((double)m[i] * (double)m[i] - 1.0 - log(v * v));
    }
  }
  s = block_sum<double, 256>(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// gm_i = t_i - m_i/num_data ;  gLs_ij = 2*H[j][i] - (Ls_ij - [i==j]/Ls_ii)/num_data for j <= i, 0 above
template <typename T>
__global__ void var_grads_kernel(const T* __restrict__ H, int64_t ldh, const T* __restrict__ Ls, int64_t ldl,
                                 const T* __restrict__ t, const T* __restrict__ m, int Mq, double inv_nd,
                                 T* __restrict__ gm, T* __restrict__ gLs, int64_t ldg) {
  __shared__ T tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x, tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  // tile of H^T: rows bi*32.., cols bj*32..  = H[bj*32 + c][bi*32 + r]
  for (int r = ty; r < 32; r += 8) {
    const int hi = bj * 32 + r, hj = bi * 32 + tx;
    tile[r][tx] = (hi < Mq && hj < Mq) ? H[(int64_t)hi * ldh + hj] : T(0);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    if (i < Mq && j < Mq) {
      T g = 0;
      if (j <= i) {
        double gg = 2.0 * (double)tile[tx][r];
        if (inv_nd != 0.0) {
          const double l = (double)Ls[(int64_t)i * ldl + j];
          gg -= inv_nd * (l - (i == j ? 1.0 / l : 0.0));
        }
        g = (T)gg;
      }
      gLs[(int64_t)i * ldg + j] = g;
    }
  }
  if (bj == 0 && ty == 0) {
    const int i = bi * 32 + tx;
    if (i < Mq) gm[i] = (T)((double)t[i] - inv_nd * (double)m[i]);
  }
}

// ================================================================================================ host side
template <typename T>
int hyp_from_raw(const T* raw_ell, const T* raw_os, const T* raw_noise, const T* c, double* hyp, cudaStream_t st) {
  if (!raw_ell || !hyp) return DSVGP_ERR_ARG;
  hyp_from_raw_kernel<T><<<1, 32, 0, st>>>(raw_ell, raw_os, raw_noise, c, hyp);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int pad_identity(double* A, int64_t ld, int Mq, int Mp, cudaStream_t st) {
  if (Mp <= Mq) return DSVGP_OK;
  pad_identity_kernel<<<(unsigned)ceil_div64((int64_t)Mp * Mp, 256), 256, 0, st>>>(A, ld, Mq, Mp);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename S, typename D>
int cast2d(const S* src, int64_t lds, D* dst, int64_t ldd, int rows, int cols, int tril, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(cols, 256), rows);
  cast2d_kernel<S, D><<<grid, 256, 0, st>>>(src, lds, dst, ldd, rows, cols, tril);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int mirror_lower(T* A, int64_t ld, int n, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(n, 32), ceil_div(n, 32)), block(32, 8);
  mirror_lower_kernel<T><<<grid, block, 0, st>>>(A, ld, n);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int add_outer(T* A, int64_t ld, int n, const T* u, const T* v, double alpha, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(n, 256), n);
  add_outer_kernel<T><<<grid, 256, 0, st>>>(A, ld, n, u, v, (T)alpha);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int tril_minus_eye(const T* Ls, int64_t ldl, T* E, int64_t lde, int n, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(n, 256), n);
  tril_minus_eye_kernel<T><<<grid, 256, 0, st>>>(Ls, ldl, E, lde, n);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int sym_phi(const double* Y, int64_t ldy, double* P, int64_t ldp, int n, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(n, 256), n);
  sym_phi_kernel<<<grid, 256, 0, st>>>(Y, ldy, P, ldp, n);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int phi_outer(const T* X, int64_t ldx, const T* u, const T* v, double* P, int64_t ldp, int n, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(ceil_div(n, 2), 128), n);
  phi_outer_kernel<T><<<grid, 128, 0, st>>>(X, ldx, u, v, P, ldp, n);
  CHECK_LAUNCH();
  return DSVGP_OK;
}
template int phi_outer<float>(const float*, int64_t, const float*, const float*, double*, int64_t, int, cudaStream_t);
template int phi_outer<double>(const double*, int64_t, const double*, const double*, double*, int64_t, int, cudaStream_t);

int phi_lower(const double* Y, int64_t ldy, double* P, int64_t ldp, int n, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(n, 256), n);
  phi_lower_kernel<<<grid, 256, 0, st>>>(Y, ldy, P, ldp, n);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int symmetrize(double* A, int64_t ld, int n, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(n, 32), ceil_div(n, 32)), block(32, 8);
  symmetrize_kernel<<<grid, block, 0, st>>>(A, ld, n);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int reduce_slabs(int rows, int cols) {
  // enough row slabs that (column blocks x slabs) covers ~4 CTAs per SM
  const int cb = ceil_div(cols, 64);
  int s = ceil_div(148 * 4, cb);
  if (s < 1) s = 1;
  const int maxs = ceil_div(rows, 64);
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}

template <typename T>
int col_dots(const T* A, const T* C, const T* B, int64_t ld, int rows, int nq, const T* m, T* pm, T* pv, int nslab,
             unsigned* cmax_bits, cudaStream_t st) {
  if (rows <= 0 || nq <= 0) return DSVGP_OK;
  if (nslab < 1) return DSVGP_ERR_ARG;
  const int rps = ceil_div(rows, nslab);
  dim3 grid(ceil_div(nq, 64), nslab);
  if (B) col_sqdiff_kernel<T><<<grid, 256, 0, st>>>(A, B, ld, rows, nq, m, rps, pm, pv);
  else col_dots_kernel<T><<<grid, 256, 0, st>>>(A, C, ld, rows, nq, m, rps, pm, pv, cmax_bits);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int col_dots_half(const void* Ah, const void* Al, int64_t ldh, const float* a_scale, const float* C, int64_t ld, int rows, int nq,
                  const float* m, float* pm, float* pv, int nslab, unsigned* cmax_bits, cudaStream_t st) {
  if (rows <= 0 || nq <= 0) return DSVGP_OK;
  if (nslab < 1 || !Ah || !Al || !a_scale || !C || (ld & 1) || (ldh & 1) || (reinterpret_cast<uintptr_t>(C) & 7) ||
      ((reinterpret_cast<uintptr_t>(Ah) | reinterpret_cast<uintptr_t>(Al)) & 3))
    return DSVGP_ERR_ARG;
  const int rps = ceil_div(rows, nslab);
  dim3 grid(ceil_div(nq, 64), nslab);
  col_dots_h_kernel<<<grid, 256, 0, st>>>(static_cast<const __half*>(Ah), static_cast<const __half*>(Al), ldh, a_scale, C, ld, rows,
                                          nq, m, rps, pm, pv, cmax_bits);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int predict_finish(const T* pm, const T* pv, int nslab, int nq, int p2, const double* hyp, double pred_jitter,
                   int add_noise, double min_var, T* mu, T* var, cudaStream_t st) {
  if (nq <= 0) return DSVGP_OK;
  predict_finish_kernel<T><<<ceil_div(nq, 256), 256, 0, st>>>(pm, pv, nslab, nq, p2, hyp, pred_jitter, add_noise,
                                                              min_var, mu, var);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

// sc[0] += data term, sc[1] += explicit dELBO/dsigma^2 ; ws: 2*ceil(nq/256) doubles
template <typename T>
int elbo_terms(const T* mu, const T* var, const T* y, int nq, const double* hyp, double w, double min_var, T* gmu,
               T* gvar, double* sc, double* ws, cudaStream_t st) {
  if (nq <= 0) return DSVGP_OK;
  const int nb = ceil_div(nq, 256);
  elbo_terms_kernel<T><<<nb, 256, 0, st>>>(mu, var, y, nq, hyp, w, min_var, gmu, gvar, ws);
  CHECK_LAUNCH();
  sum_scalar_parts_kernel<<<1, 256, 0, st>>>(ws, nb, 2, 2, sc);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int pll_terms(const T* mu, const T* var, const T* y, int nq, double w, double min_var, T* gmu, T* gvar, double* sc,
              double* ws, cudaStream_t st) {
  if (nq <= 0) return DSVGP_OK;
  const int nb = ceil_div(nq, 256);
  pll_terms_kernel<T><<<nb, 256, 0, st>>>(mu, var, y, nq, w, min_var, gmu, gvar, ws);
  CHECK_LAUNCH();
  sum_scalar_parts_kernel<<<1, 256, 0, st>>>(ws, nb, 2, 2, sc);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

// gsc[0..3] += {d ell, d os, d sigma^2, d c} ; ws: 4*blocks doubles (blocks <= 296)
template <typename T>
int pred_bwd_scalars(const T* gmu, const T* gvar, int nq, int p2, const double* hyp, int add_noise, double* gsc,
                     double* ws, cudaStream_t st) {
  if (nq <= 0) return DSVGP_OK;
  int nb = ceil_div(nq, 256);
  if (nb > 296) nb = 296;
  pred_bwd_scalars_kernel<T><<<nb, 256, 0, st>>>(gmu, gvar, nq, p2, hyp, add_noise, ws);
  CHECK_LAUNCH();
  sum_scalar_parts_kernel<<<1, 256, 0, st>>>(ws, nb, 4, 4, gsc);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int dA_apply(const T* A, T* C, T* Ag, int64_t ld, int rows, int nq, const T* m, const T* gmu, const T* gvar, T* tp,
             int nslab, T* t, T* Clo, T* Aglo, cudaStream_t st) {
  if (rows <= 0 || nq <= 0) return DSVGP_OK;
  const int cps = ceil_div(ceil_div(nq, nslab), 32) * 32;
  const int ns = ceil_div(nq, cps);
  dim3 grid(ceil_div(rows, 8), ns);
  dA_kernel<T><<<grid, 256, 0, st>>>(A, C, Ag, ld, rows, nq, m, gmu, gvar, cps, tp, Clo, Aglo);
  CHECK_LAUNCH();
  sum_parts_kernel<T><<<ceil_div(rows, 256), 256, 0, st>>>(tp, ns, rows, t);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int dA_apply_half(const float* A, const float* C, int64_t ld, int rows, int nq, const float* m, const float* gmu,
                  const float* gvar, float* tp, int nslab, float* t, void* dAh, void* dAl, void* Agh, void* Agl, int64_t ldh,
                  const float* s_dA, const float* s_Ag, const void* Ah, const void* Al, const float* a_scale, cudaStream_t st) {
  if (rows <= 0 || nq <= 0) return DSVGP_OK;
  if (!A && (!Ah || !Al || !a_scale)) return DSVGP_ERR_ARG;
  const int cps = ceil_div(ceil_div(nq, nslab), 32) * 32;
  const int ns = ceil_div(nq, cps);
  dim3 grid(ceil_div(rows, 8), ns);
  dA_half_kernel<<<grid, 256, 0, st>>>(A, C, ld, rows, nq, m, gmu, gvar, cps, tp, static_cast<__half*>(dAh),
                                       static_cast<__half*>(dAl), static_cast<__half*>(Agh), static_cast<__half*>(Agl), ldh,
                                       s_dA, s_Ag, static_cast<const __half*>(Ah), static_cast<const __half*>(Al), a_scale);
  CHECK_LAUNCH();
  sum_parts_kernel<float><<<ceil_div(rows, 256), 256, 0, st>>>(tp, ns, rows, t);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

// out[0] += KL ; ws: 296 doubles
template <typename T>
int kl_divergence(const T* m, const T* Ls, int64_t ld, int Mq, double* out, double* ws, cudaStream_t st) {
  if (Mq <= 0) return DSVGP_OK;
  int nb = (int)ceil_div64((int64_t)Mq * Mq, 256 * 8);
  if (nb > 296) nb = 296;
  kl_kernel<T><<<nb, 256, 0, st>>>(m, Ls, ld, Mq, ws);
  CHECK_LAUNCH();
  sum_scalar_parts_kernel<<<1, 256, 0, st>>>(ws, nb, 1, 1, out);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template <typename T>
int var_grads(const T* H, int64_t ldh, const T* Ls, int64_t ldl, const T* t, const T* m, int Mq, double inv_nd, T* gm,
              T* gLs, int64_t ldg, cudaStream_t st) {
  if (Mq <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(Mq, 32), ceil_div(Mq, 32)), block(32, 8);
  var_grads_kernel<T><<<grid, block, 0, st>>>(H, ldh, Ls, ldl, t, m, Mq, inv_nd, gm, gLs, ldg);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

// All the small gradients of a step in ONE launch (they used to be ~20 one-microsecond torch kernels at the very end of the step,
// launch-bound: 0.1 ms): out = [dZ (nZ) | dV_z (nV) | d c | d raw_outputscale | d raw_lengthscale | d raw_noise] in the model dtype
// from the fp64 buffer small = [scalars(8) | dZ | dV_z]; chain rule of the softplus transforms from hyp[4..6];
// noise_mode 1: d noise = scalars[1] + scalars[6] (ELBO / PLL step), 0: scalars[6] (generic predictive backward).
template <typename T>
__global__ void collect_grads_kernel(const double* __restrict__ small, int nZ, int nV, const double* __restrict__ hyp,
                                     int noise_mode, T* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, n = nZ + nV;
  if (i < n) out[i] = (T)small[8 + i];
  else if (i == n) out[i] = (T)small[7];
  else if (i == n + 1) out[i] = (T)(small[5] * hyp[5]);
  else if (i == n + 2) out[i] = (T)(small[4] * hyp[4]);
  else if (i == n + 3) out[i] = (T)(((noise_mode ? small[1] : 0.0) + small[6]) * hyp[6]);
}

template <typename T>
int collect_grads(const double* small, int nZ, int nV, const double* hyp, int noise_mode, T* out, cudaStream_t st) {
  if (!small || !hyp || !out || nZ < 0 || nV < 0) return DSVGP_ERR_ARG;
  collect_grads_kernel<T><<<ceil_div(nZ + nV + 4, 256), 256, 0, st>>>(small, nZ, nV, hyp, noise_mode, out);
  CHECK_LAUNCH();
  return DSVGP_OK;
}
template int collect_grads<float>(const double*, int, int, const double*, int, float*, cudaStream_t);
template int collect_grads<double>(const double*, int, int, const double*, int, double*, cudaStream_t);

// ------------------------------------------------------------------------------------ fp64 tensor peak (measurement)
// Register-resident DMMA m8n8k4 chains, 8 independent accumulators per warp, no memory traffic: what the fp64 tensor pipe
// sustains on this device.  bench.py times it with CUDA events and uses the result as the roofline denominator of every
// fp64 product (MEASURED_PEAKS.json has no fp64 figure).
__global__ void __launch_bounds__(256)
dmma_peak_kernel(double* __restrict__ out, int iters) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  if (s == -1.0) out[0] = s;        // never true: keeps the chains alive
}

int dmma_peak(int iters, int ctas, double* out, double* flops_host, cudaStream_t st) {
  if (iters <= 0 || ctas <= 0 || !out) return DSVGP_ERR_ARG;
  dmma_peak_kernel<<<ctas, 256, 0, st>>>(out, iters);
  CHECK_LAUNCH();
  if (flops_host) *flops_host = (double)ctas * 8.0 /*warps*/ * (double)iters * 8.0 /*chains*/ * 512.0 /*2*8*8*4*/;
  return DSVGP_OK;
}

#define INST(T)                                                                                                    \
  template int hyp_from_raw<T>(const T*, const T*, const T*, const T*, double*, cudaStream_t);                     \
  template int mirror_lower<T>(T*, int64_t, int, cudaStream_t);                                                    \
  template int tril_minus_eye<T>(const T*, int64_t, T*, int64_t, int, cudaStream_t);                                                    \
  template int add_outer<T>(T*, int64_t, int, const T*, const T*, double, cudaStream_t);                           \
  template int col_dots<T>(const T*, const T*, const T*, int64_t, int, int, const T*, T*, T*, int, unsigned*, cudaStream_t);  \
  template int predict_finish<T>(const T*, const T*, int, int, int, const double*, double, int, double, T*, T*,    \
                                 cudaStream_t);                                                                    \
  template int elbo_terms<T>(const T*, const T*, const T*, int, const double*, double, double, T*, T*, double*,    \
                             double*, cudaStream_t);                                                               \
  template int pll_terms<T>(const T*, const T*, const T*, int, double, double, T*, T*, double*, double*,           \
                            cudaStream_t);                                                                         \
  template int pred_bwd_scalars<T>(const T*, const T*, int, int, const double*, int, double*, double*,             \
                                   cudaStream_t);                                                                  \
  template int dA_apply<T>(const T*, T*, T*, int64_t, int, int, const T*, const T*, const T*, T*, int, T*, T*, T*, \
                           cudaStream_t);                                                                          \
  template int kl_divergence<T>(const T*, const T*, int64_t, int, double*, double*, cudaStream_t);                 \
  template int var_grads<T>(const T*, int64_t, const T*, int64_t, const T*, const T*, int, double, T*, T*,         \
                            int64_t, cudaStream_t);
INST(float)
INST(double)
#undef INST
template int cast2d<double, float>(const double*, int64_t, float*, int64_t, int, int, int, cudaStream_t);
template int cast2d<float, double>(const float*, int64_t, double*, int64_t, int, int, int, cudaStream_t);
template int cast2d<double, double>(const double*, int64_t, double*, int64_t, int, int, int, cudaStream_t);
template int cast2d<float, float>(const float*, int64_t, float*, int64_t, int, int, int, cudaStream_t);

}  // namespace dsvgp
