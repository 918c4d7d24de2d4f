// Blocked right-looking Cholesky of the jittered K_zz in fp64 plus the explicit inverse of the factor.
//
// Replaces psd_safe_cholesky(K_zz.double()) (DirectionalGradVariationalStrategy.py:72-75, cuSOLVER potrf in the
// reference) and turns both triangular solves (:181,:183) into one triangular matrix product W * K_zx with
// W = L^-1, which is what lets the whitening run as a plain tiled tensor-core GEMM.
//
// The matrix is padded with an identity block to Mp = nb0 * 2^nlev so that every level of the recursive
// inverse has uniform blocks:  for a 2x2 block-lower L = [[L11,0],[L21,L22]],
//     W = [[W11, 0], [-W22 * (L21 * W11), W22]],
// applied bottom-up; each level is two batched GEMMs.  Diagonal nb0 x nb0 blocks are factorised AND inverted
// inside one CTA in shared memory.  Status goes to a device int (0 ok, i+1 = first non-positive pivot), never to
// the host: the caller decides when to look (no hidden synchronisation).
#include "chol.cuh"
#include "common.cuh"
#include "gemm.cuh"

namespace dsvgp {

constexpr int POTRF_THREADS = 512;

void chol_plan(int Mq, int* Mp, int* nb0, int* nlev) {
  int k = 0;
  while (ceil_div(Mq, 1 << k) > 112) ++k;
  int b = ceil_div(Mq, 1 << k);
  b = ceil_div(b, 4) * 4;
  *Mp = b << k;
  *nb0 = b;
  *nlev = k;
}

// factorise the nb x nb block at A (lower part read), write L block (upper zeroed) and its inverse W block
__global__ void __launch_bounds__(POTRF_THREADS)
potrf_inv_block(const double* __restrict__ A, int64_t lda, double* __restrict__ L, int64_t ldl,
                double* __restrict__ W, int64_t ldw, int nb, int* __restrict__ info, int row_offset) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = nb + 1;
  double* Ls = reinterpret_cast<double*>(smem_raw);   // [nb][ld]
  double* Ws = Ls + nb * ld;                          // [nb][ld]
  double* dg = Ws + nb * ld;                          // [nb]
  const int tid = threadIdx.x;
  for (int e = tid; e < nb * nb; e += POTRF_THREADS) {
    const int i = e / nb, c = e % nb;
    Ls[i * ld + c] = (c <= i) ? A[(int64_t)i * lda + c] : 0.0;
    Ws[i * ld + c] = 0.0;
  }
  for (int j = 0; j < nb; ++j) {
    __syncthreads();
    const double djj = Ls[j * ld + j];
    double ljj;
    if (!(djj > 0.0)) {          // also catches NaN
      if (tid == 0) atomicCAS(info, 0, row_offset + j + 1);
      ljj = nan("");
    } else {
      ljj = sqrt(djj);
    }
    if (tid == 0) dg[j] = ljj;
    const double inv = 1.0 / ljj;
    for (int i = j + 1 + tid; i < nb; i += POTRF_THREADS) Ls[i * ld + j] *= inv;
    __syncthreads();
    const int m = nb - j - 1;
    for (int e = tid; e < m * m; e += POTRF_THREADS) {
      const int i = j + 1 + e / m, c = j + 1 + e % m;
      if (c <= i) Ls[i * ld + c] -= Ls[i * ld + j] * Ls[c * ld + j];
    }
  }
  __syncthreads();
  for (int j = tid; j < nb; j += POTRF_THREADS) Ls[j * ld + j] = dg[j];
  __syncthreads();
  // inverse by forward substitution, 4 threads per column (k-sum split 4 ways, combined by shuffles)
  {
    const int col = tid >> 2, part = tid & 3;
    for (int i = 0; i < nb; ++i) {
      double s = 0.0;
      if (col < nb && col <= i)
        for (int k = col + part; k < i; k += 4) s += Ls[i * ld + k] * Ws[k * ld + col];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0 && col < nb && col <= i) Ws[i * ld + col] = ((i == col ? 1.0 : 0.0) - s) / Ls[i * ld + i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < nb * nb; e += POTRF_THREADS) {
    const int i = e / nb, c = e % nb;
    L[(int64_t)i * ldl + c] = Ls[i * ld + c];
    W[(int64_t)i * ldw + c] = Ws[i * ld + c];
  }
}

int chol_factor_inverse(double* Awork, int64_t lda, double* L, int64_t ldl, double* W, int64_t ldw, int Mp, int nb0,
                        int nlev, int* info, cudaStream_t st) {
  if (Mp <= 0) return DSVGP_OK;
  if (!Awork || !L || !W || !info || nb0 <= 0 || nb0 > 128 || (nb0 << nlev) != Mp) return DSVGP_ERR_ARG;
  const int nblk = 1 << nlev;
  const size_t smem = sizeof(double) * (size_t)(2 * nb0 * (nb0 + 1) + nb0);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(potrf_inv_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemsetAsync(info, 0, sizeof(int), st);
  for (int k = 0; k < nblk; ++k) {
    const int64_t o = (int64_t)k * nb0;
    potrf_inv_block<<<1, POTRF_THREADS, smem, st>>>(Awork + o * lda + o, lda, L + o * ldl + o, ldl, W + o * ldw + o,
                                                    ldw, nb0, info, (int)o);
    CHECK_LAUNCH();
    const int m = Mp - (int)o - nb0;
    if (m > 0) {
      double* L21 = L + (o + nb0) * ldl + o;
      // panel  L21 = A21 * W11^T   (op(B) = W11^T is upper triangular)
      int rc = gemm1<double>(false, true, m, nb0, nb0, 1.0, Awork + (o + nb0) * lda + o, lda, W + o * ldw + o, ldw, 0.0,
                             L21, ldl, TRI_NONE, TRI_UPPER, 0, st);
      if (rc) return rc;
      // trailing update  A22 -= L21 * L21^T  (lower tiles only)
      rc = gemm1<double>(false, true, m, m, nb0, -1.0, L21, ldl, L21, ldl, 1.0,
                         Awork + (o + nb0) * lda + (o + nb0), lda, TRI_NONE, TRI_NONE, 1, st);
      if (rc) return rc;
    }
  }
  // recursive inverse; Awork (no longer needed) is the scratch for T = L21 * W11
  for (int lev = 0; lev < nlev; ++lev) {
    const int b = nb0 << lev, npairs = nblk >> (lev + 1);
    const int64_t sL = (int64_t)2 * b * ldl + 2 * b, sW = (int64_t)2 * b * ldw + 2 * b, sT = (int64_t)2 * b * lda + 2 * b;
    int rc = gemm<double>(false, false, b, b, b, 1.0, L + (int64_t)b * ldl, ldl, W, ldw, 0.0, Awork + (int64_t)b * lda,
                          lda, TRI_NONE, TRI_LOWER, 0, npairs, sL, sW, sT, st);
    if (rc) return rc;
    rc = gemm<double>(false, false, b, b, b, -1.0, W + (int64_t)b * ldw + b, ldw, Awork + (int64_t)b * lda, lda, 0.0,
                      W + (int64_t)b * ldw, ldw, TRI_LOWER, TRI_NONE, 0, npairs, sW, sT, sW, st);
    if (rc) return rc;
  }
  return DSVGP_OK;
}

}  // namespace dsvgp
