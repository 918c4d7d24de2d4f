// Scales and operand splits for the 3xFP16 tensor-core products (definitions in tc_prep.cu).
#pragma once
#include "common.cuh"

namespace dsvgp {

// *out = max(*out, max |x_ij|) as the bit pattern of a non-negative float (zero it first).  mode 2: entries of tril(x) - I.
template <typename T>
int absmax(const T* x, int64_t ld, int rows, int cols, int mode, unsigned* out, cudaStream_t st);

// power-of-two scales of every operand of the step from a-priori bounds; layout in tc_prep.cu
int tc_scales(const double* hyp, double jitter, const unsigned* maxbits, int Mq, float* scales, int stage, cudaStream_t st);

// (hi, lo) halves of op(src) * *scale [and of its transpose]; mode 0 as is, 1 tril, 2 tril - I
template <typename S>
int split_half(const S* src, int64_t lds, int rows, int cols, int mode, const float* scale, void* hi, void* lo, int64_t ldh,
               void* hiT, void* loT, int64_t ldhT, cudaStream_t st);

// (hi, lo) halves of (E + E^T + E E^T) * *scale from the lower triangles of E and P = E E^T (n x n, fp32)
int build_d_split(const float* E, int64_t lde, const float* P, int64_t ldp, int n, const float* scale, void* hi, void* lo,
                  int64_t ldh, cudaStream_t st);

// *out_bits = max(*out_bits, max |E + E^T + E E^T|) (bit pattern of a non-negative float): the measured bound of D
int build_d_absmax(const float* E, int64_t lde, const float* P, int64_t ldp, int n, unsigned* out_bits, cudaStream_t st);

}  // namespace dsvgp
