// Minibatch gather + select_cols_of_y (definitions in data.cu).
#pragma once
#include "common.cuh"

namespace dsvgp {

// X: N x d, Y: N x ycols (row-major, contiguous); idx: n int64 row indices (device) or null for rows 0..n-1;
// cols_host: p+1 column indices of Y (host).  Outputs xb (n x d), yb (n*(p+1), interleaved), V (n*p x d one-hot, or null).
template <typename T>
int gather_batch(const T* X, const T* Y, int64_t N, int d, int ycols, const int64_t* idx, int n, int p,
                 const int* cols_host, T* xb, T* yb, T* V, cudaStream_t st);

}  // namespace dsvgp
