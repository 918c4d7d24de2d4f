// Input side of a training step: minibatch gather + select_cols_of_y in one pass.
//
// The reference draws a shuffled minibatch through a DataLoader, moves it to the GPU, keeps the function-value column
// and `minibatch_dim` randomly chosen gradient columns of y, builds the matching one-hot directions and repeats them
// for every point (directionalvi/directional_vi.py:68-90, :229-241) -- a host loop plus ~6 small device ops and an
// H2D copy per step.  Here the dataset stays resident in HBM and one kernel writes, for minibatch row i = idx[i]:
//   xb[i, :]            = X[idx[i], :]
//   yb[i*(p+1) + a]     = Y[idx[i], cols[a]]                 (interleaved [f, d_1..d_p], :241 reshape)
//   V[(i*p + b), :]     = e_{cols[1+b]-1}                    (canonical direction rows, :87-88 + .repeat(n,1))
// `cols` (p+1 sorted column indices, cols[0] = 0) is chosen on the host exactly as the reference does (Python's
// random.sample) and travels in the kernel parameters.
#include "common.cuh"
#include "data.cuh"

namespace dsvgp {

struct ColSel { int c[DSVGP_MAXP + 1]; };

template <typename T>
__global__ void __launch_bounds__(256)
gather_batch_kernel(const T* __restrict__ X, const T* __restrict__ Y, int d, int ycols, const int64_t* __restrict__ idx,
                    int n, int p, const __grid_constant__ ColSel cols, T* __restrict__ xb, T* __restrict__ yb,
                    T* __restrict__ V) {
  const int per = d + (p + 1) + (V ? p * d : 0);                 // output elements per minibatch point
  const int64_t total = (int64_t)n * per;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int i = (int)(e / per), r = (int)(e - (int64_t)i * per);
    const int64_t src = idx ? idx[i] : i;
    if (r < d) {
      xb[(int64_t)i * d + r] = X[src * d + r];
    } else if (r < d + p + 1) {
      const int a = r - d;
      yb[(int64_t)i * (p + 1) + a] = Y[src * ycols + cols.c[a]];
    } else {
      const int q = r - d - (p + 1), b = q / d, c = q - b * d;
      V[((int64_t)i * p + b) * d + c] = (c == cols.c[1 + b] - 1) ? T(1) : T(0);
    }
  }
}

template <typename T>
int gather_batch(const T* X, const T* Y, int64_t N, int d, int ycols, const int64_t* idx, int n, int p,
                 const int* cols_host, T* xb, T* yb, T* V, cudaStream_t st) {
  if (n <= 0) return DSVGP_OK;
  if (!X || !Y || !cols_host || !xb || !yb || d <= 0 || p < 0 || p > DSVGP_MAXP || ycols < 1 || N < 0) return DSVGP_ERR_ARG;
  ColSel cs;
  for (int a = 0; a <= p; ++a) {
    if (cols_host[a] < 0 || cols_host[a] >= ycols) return DSVGP_ERR_ARG;
    if (a > 0 && V && (cols_host[a] < 1 || cols_host[a] > d)) return DSVGP_ERR_ARG;
    cs.c[a] = cols_host[a];
  }
  const int per = d + (p + 1) + (V ? p * d : 0);
  int64_t blocks = ceil_div64((int64_t)n * per, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_batch_kernel<T><<<(int)blocks, 256, 0, st>>>(X, Y, d, ycols, idx, n, p, cs, xb, yb, V);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

template int gather_batch<float>(const float*, const float*, int64_t, int, int, const int64_t*, int, int, const int*,
                                 float*, float*, float*, cudaStream_t);
template int gather_batch<double>(const double*, const double*, int64_t, int, int, const int64_t*, int, int, const int*,
                                  double*, double*, double*, cudaStream_t);

}  // namespace dsvgp
