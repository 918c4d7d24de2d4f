// Triangle-aware batched GEMM (definitions in gemm.cu).
#pragma once
#include "common.cuh"

namespace dsvgp {

enum { TRI_NONE = 0, TRI_LOWER = 1, TRI_UPPER = 2 };

// C = alpha*op(A)*op(B) + beta*C, row-major.  a_tri/b_tri: structure of op(A)/op(B) as product operands.
// c_tri = 1: only tiles touching the lower triangle of C are computed (the rest of C is left untouched).
// batch > 1: operand z lives at base + z*stride (element strides sA, sB, sC).
template <typename T>
int gemm(bool ta, bool tb, int M, int N, int K, T alpha, const T* A, int64_t lda, const T* B, int64_t ldb, T beta,
         T* C, int64_t ldc, int a_tri, int b_tri, int c_tri, int batch, int64_t sA, int64_t sB, int64_t sC,
         cudaStream_t st, const T* D = nullptr, int64_t ldd = 0,    // D != null: C = alpha*AB + beta*D
         T* C2 = nullptr, int64_t ldc2 = 0, const T* D2 = nullptr, int64_t ldd2 = 0,    // C2 = C + D2
         int c_off = 0);   // c_tri == 1 with c_off > 0: lower trapezoid, tiles up to c_off columns right of the diagonal

template <typename T>
inline int gemm1(bool ta, bool tb, int M, int N, int K, T alpha, const T* A, int64_t lda, const T* B, int64_t ldb,
                 T beta, T* C, int64_t ldc, int a_tri, int b_tri, int c_tri, cudaStream_t st) {
  return gemm<T>(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, a_tri, b_tri, c_tri, 1, 0, 0, 0, st);
}

// fp64 products: 1 (default) = cp.async-pipelined kernel, 0 = the register-staged kernel of round 1 (A/B knob)
void set_gemm64_async(int on);
int get_gemm64_async();
// fp64  C = alpha A B^T + beta C  with K <= 128 (the rank-96 Cholesky updates): 1 = whole-K-in-shared kernel, 0 (default: the former measured no gain) = general kernel
void set_rank_update(int on);
int get_rank_update();

}  // namespace dsvgp
