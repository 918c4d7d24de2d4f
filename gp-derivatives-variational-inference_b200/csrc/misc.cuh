// Elementwise / reduction kernels (definitions in misc.cu).
#pragma once
#include "common.cuh"

namespace dsvgp {

template <typename T>
int hyp_from_raw(const T* raw_ell, const T* raw_os, const T* raw_noise, const T* c, double* hyp, cudaStream_t st);
int pad_identity(double* A, int64_t ld, int Mq, int Mp, cudaStream_t st);
template <typename S, typename D>
int cast2d(const S* src, int64_t lds, D* dst, int64_t ldd, int rows, int cols, int tril, cudaStream_t st);
template <typename T> int mirror_lower(T* A, int64_t ld, int n, cudaStream_t st);
template <typename T> int add_outer(T* A, int64_t ld, int n, const T* u, const T* v, double alpha, cudaStream_t st);
template <typename T> int tril_minus_eye(const T* Ls, int64_t ldl, T* E, int64_t lde, int n, cudaStream_t st);
int sym_phi(const double* Y, int64_t ldy, double* P, int64_t ldp, int n, cudaStream_t st);
int phi_lower(const double* Y, int64_t ldy, double* P, int64_t ldp, int n, cudaStream_t st);
template <typename T>
int phi_outer(const T* X, int64_t ldx, const T* u, const T* v, double* P, int64_t ldp, int n, cudaStream_t st);
int symmetrize(double* A, int64_t ld, int n, cudaStream_t st);
int reduce_slabs(int rows, int cols);
template <typename T>
int col_dots(const T* A, const T* C, const T* B, int64_t ld, int rows, int nq, const T* m, T* pm, T* pv, int nslab,
             unsigned* cmax_bits, cudaStream_t st);
template <typename T>
int predict_finish(const T* pm, const T* pv, int nslab, int nq, int p2, const double* hyp, double pred_jitter,
                   int add_noise, double min_var, T* mu, T* var, cudaStream_t st);
template <typename T>
int elbo_terms(const T* mu, const T* var, const T* y, int nq, const double* hyp, double w, double min_var, T* gmu,
               T* gvar, double* sc, double* ws, cudaStream_t st);
int dA_apply_half(const float* A, const float* C, int64_t ld, int rows, int nq, const float* m, const float* gmu,
                  const float* gvar, float* tp, int nslab, float* t, void* dAh, void* dAl, void* Agh, void* Agl, int64_t ldh,
                  const float* s_dA, const float* s_Ag, const void* Ah, const void* Al, const float* a_scale, cudaStream_t st);
// A == nullptr: A is read as the two-half split (Ah, Al; leading dimension ldh) of A * *a_scale
// col_dots with A given only as that split (3xFP16 training step: the whitening product writes no fp32 A)
int col_dots_half(const void* Ah, const void* Al, int64_t ldh, const float* a_scale, const float* C, int64_t ld, int rows, int nq,
                  const float* m, float* pm, float* pv, int nslab, unsigned* cmax_bits, cudaStream_t st);
template <typename T>
int pll_terms(const T* mu, const T* var, const T* y, int nq, double w, double min_var, T* gmu, T* gvar, double* sc,
              double* ws, cudaStream_t st);
template <typename T>
int pred_bwd_scalars(const T* gmu, const T* gvar, int nq, int p2, const double* hyp, int add_noise, double* gsc,
                     double* ws, cudaStream_t st);
template <typename T>
int dA_apply(const T* A, T* C, T* Ag, int64_t ld, int rows, int nq, const T* m, const T* gmu, const T* gvar, T* tp,
             int nslab, T* t, T* Clo, T* Aglo, cudaStream_t st);
template <typename T>
int kl_divergence(const T* m, const T* Ls, int64_t ld, int Mq, double* out, double* ws, cudaStream_t st);
template <typename T>
int var_grads(const T* H, int64_t ldh, const T* Ls, int64_t ldl, const T* t, const T* m, int Mq, double inv_nd, T* gm,
              T* gLs, int64_t ldg, cudaStream_t st);

int dmma_peak(int iters, int ctas, double* out, double* flops_host, cudaStream_t st);

// every small gradient of a step (dZ, dV_z, d c, d raw_os, d raw_ell, d raw_noise) in the model dtype, one launch
template <typename T>
int collect_grads(const double* small, int nZ, int nV, const double* hyp, int noise_mode, T* out, cudaStream_t st);

}  // namespace dsvgp
