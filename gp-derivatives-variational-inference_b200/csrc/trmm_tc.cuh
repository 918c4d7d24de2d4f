// tcgen05 / TMA 3xTF32 GEMM (definitions in trmm_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace dsvgp {

// 1 if the operands satisfy the TMA constraints (16-byte aligned bases, leading dims % 4 == 0, for a K x N
// row-major B the last 32-column block inside the row) and the driver exposes cuTensorMapEncodeTiled.
int gemm_tc_supported(const float* A, int64_t lda, const float* B, int64_t ldb, int b_kmajor, int N);

// C = alpha * A * B + beta * D (+ C2 = C + D2).  A: M x K row-major (zeros outside its triangle).  B: K x N row-major
// (b_kmajor = 0) or N x K row-major (b_kmajor = 1).  *_lo = x - trunc_tf32(x) (split_lo).  a_tri: 1 lower / 2 upper
// k-range trimming; c_lower: skip tiles above the diagonal; chunk: k-blocks (of 32) per tensor-core accumulation chain.
int gemm_tc(const float* Ah, const float* Al, int64_t lda, const float* Bh, const float* Bl, int64_t ldb, int b_kmajor,
            int M, int N, int K, float alpha, float beta, float* C, int64_t ldc, const float* D, int64_t ldd, float* C2,
            int64_t ldc2, const float* D2, int64_t ldd2, int a_tri, int c_lower, int chunk, float* Clo, float* C2lo,
            int nsplit, float* split_ws, cudaStream_t st);
// Clo / C2lo (optional): lo parts of C / C2 written by the same epilogue (operands of a following product).
// nsplit > 1: split-K over gridDim.z; raw partial sums go to split_ws (nsplit * M * round_up(N,4) floats) and are
// summed in fp64 by a second kernel (needs C2 == Clo == null).

// 3xFP16 variant (tcgen05 kind::f16): operands are the two-half splits (hi, lo) of x * s with a power-of-two scale s per
// matrix (fp16 has the 11 significand bits of tf32 but 5 exponent bits; the scale comes from an a-priori bound so that
// nothing overflows).  Same pipeline, twice the K per shared-memory byte and per MMA instruction.  *ab_inv (device) =
// 1 / (sA * sB).  Outputs: any of C (fp32), C2 = C + D2 (fp32), and the two-half splits Ch/Cl of C * *c_scale and
// C2h/C2l of C2 * *c2_scale for a following product.  Leading dimensions of half arrays % 8 == 0; a K x N row-major B
// needs ldb >= round_up(N, 64).  Pointers to half arrays are passed as void*.
int gemm_tch_supported(const void* A, int64_t lda, const void* B, int64_t ldb, int b_kmajor, int N);
int gemm_tch(const void* Ah, const void* Al, int64_t lda, const void* Bh, const void* Bl, int64_t ldb, int b_kmajor, int M,
             int N, int K, float alpha, float beta, const float* ab_inv, float* C, int64_t ldc, const float* D, int64_t ldd,
             float* C2, int64_t ldc2, const float* D2, int64_t ldd2, void* Ch, void* Cl, int64_t ldch, const float* c_scale,
             void* C2h, void* C2l, int64_t ldc2h, const float* c2_scale, int a_tri, int c_lower, int chunk, int nsplit,
             float* split_ws, cudaStream_t st);

// 1: one CTA per 128x256 tile (tcgen05 cta_group::1); 2: CTA pairs on 256x256 tiles (cta_group::2, 2-SM TMA, multicast commit)
void set_tc_cta_group(int cg);
int get_tc_cta_group();
// 3xFP16 CTA-pair kernel: N extent of a tile, 256 (one pair per SM pair) or 128 (two pairs resident per SM pair)
void set_tc_tile_n(int n);
int get_tc_tile_n();
// CTA-pair kernels: 1 (default) = persistent pairs walking a host-balanced work list (set-up once per SM, operand ring kept
// full across tiles, the next tile's first chunks under the previous tile's store phase); 0 = one CTA pair per tile
// persistent kernels: at most n CTA pairs (0 = all the device holds) -- for a product that is to share the GPU with other work
void set_tc_max_pairs(int n);
int get_tc_max_pairs();
void set_tc_persistent(int on);
// profiling aid of the persistent kernels: when buf != null, the MMA warp and the first epilogue warp of every pair leader write
// SM clock stamps for each of their first cap_items work items: buf[8 * item + {0: MMA warp reaches the item, 1: first k-block
// issued, 2: last k-block issued, 3: epilogue reaches the item, 4: first chunk complete, 5: last chunk added, 6: tile stored}]
void set_tc_trace(long long* buf, int cap_items);
int get_tc_persistent();

int split_lo(const float* x, int64_t ldx, float* lo, int64_t ldl, int rows, int cols, cudaStream_t st);
int transpose_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, cudaStream_t st);

namespace tc {
// work lists of the persistent kernels (pure host function; see trmm_tc.cu)
std::vector<int> build_sched(int M, int N, int K, int a_tri, int c_lower, int nz, int P, int bke, int* nz_eff);
}

}  // namespace dsvgp
