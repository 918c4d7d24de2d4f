// tcgen05 / TMA 3xTF32 GEMM for the big whitening products of the fp32 model (sm_100a).
//
//   C[M x N] = alpha * A[M x K] * B + beta * D      (+ optional second output C2 = C + D2)
//
//   A : small square factor (W = L^-1, E = L_s - I, or an explicit transpose of one), row-major, K contiguous
//       -> UMMA K-major operand;  B : the big minibatch-sized matrix, either K x N row-major (N contiguous,
//       UMMA MN-major -- the four triangular products) or N x K row-major (K-major -- the Gram matrix A diag(g) A^T).
//
// fp32 accuracy on TF32 tensor cores: every operand comes as the raw fp32 array (the tensor core ignores the low
// 13 mantissa bits: that IS the "hi" part) plus a precomputed "lo" array  lo = x - trunc_tf32(x);  three MMAs per
// k-step accumulate  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.  The tensor-core accumulation chain is kept SHORT: the
// accumulator in tensor memory is double-buffered and restarted every `chunk` k-blocks, and the epilogue warps add
// each finished chunk into fp32 master accumulators in registers while the next chunk is being multiplied.  (A long
// fp32 tensor-core chain measured ~1e-6 * |A||B| of error at K = 3072, which the ill-conditioned whitening
// amplifies ~400x; see DESIGN.md "precision".)
//
// Structure (one CTA per 128 x 256 output tile, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A_hi, A_lo, B_hi, B_lo (128B swizzle) into a ring of
//               smem stages, completion on mbarriers
//   warp 1      allocates TMEM (512 columns = two 128x256 fp32 accumulators); one elected lane issues tcgen05.mma
//               kind::tf32 (M=128, N=256, K=8), tcgen05.commit releases smem stages / publishes finished chunks
//   warps 2-9   epilogue: tcgen05.ld the finished chunk (each warp: its 32-lane quarter x 128 columns), add into the
//               register master sums, release the TMEM buffer; at the end apply alpha/beta and store.
// Triangular structure of A is exploited by trimming the k-range of each row tile; `c_lower` skips tiles above the
// diagonal (Gram matrix).
// Kernels: gemm_tc*_kernel (one CTA per tile), gemm_tc*2_kernel (a CTA PAIR per 256 x 256 tile, tcgen05 cta_group::2) and -- the
// default -- gemm_tc*2p_kernel: PERSISTENT CTA pairs walking a host-built work list (section "Persistent CTA-pair variant" below:
// set-up once per SM, operand ring kept full across tiles, setmaxnreg, stream-K for split-K products, per-item clock trace).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "trmm_tc.cuh"

namespace dsvgp {

namespace tc {

constexpr int BM = 128, BN = 256, BK = 32;               // BK fp32 = 128 bytes = one swizzle row
constexpr int BKH = 64;                                  // fp16 variant: 64 halves = the same 128-byte row, twice the K
constexpr int A_BYTES = BM * BK * 4;                     // 16 KB (either element type)
constexpr int THREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;              // shared::cluster address of the same offset in the pair's CTA 0
// CG = 1: one CTA per 128x256 tile.  CG = 2: a CTA PAIR (tcgen05 cta_group::2) per 256x256 tile -- each CTA stages its
// own 128 rows of A and its own 128 columns of B, so the operand bytes per SM drop by a third and one more stage fits.
// BNT = 256 (one CTA or CTA pair per SM / SM pair) or, for CTA pairs only, BNT = 128 with TWO pairs resident per SM pair
// (half the tensor memory, half the ring, 4 epilogue warps each): while one pair is in its fixed phases -- barrier set-up,
// pipeline fill, the store phase of its tile -- the other pair's main loop keeps the tensor pipe busy.
template <int CG, int BNT = BN> struct Geo {
  static constexpr int BN_LOCAL = BNT / CG;
  static constexpr int B_BYTES = BN_LOCAL * BK * 4;                  // 32 KB / 16 KB / 8 KB
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;      // hi + lo of both operands: 96 KB / 64 KB / 48 KB
  static constexpr int STAGES = CG == 1 ? 2 : (BNT == BN ? 3 : 2);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int EPI = 4 * (BNT / 128);                        // epilogue warps: one per (lane quarter, 128 columns)
  static constexpr int NTHREADS = 32 * (2 + EPI);
  static constexpr int TMEM_COLS = 2 * BNT;                          // double-buffered accumulator
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {          // arrive on the pair leader's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA loads: executed by both CTAs of a pair into their own shared memory; the bytes are counted on CTA 0's barrier
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {           // arrive on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// UMMA shared-memory descriptor, descriptor version 1 (Blackwell).  layout_type 2 = 128B swizzle (16-byte atoms; K-major
// operands); layout_type 1 = 128B swizzle with 32-byte atoms, the ONLY layout tcgen05 accepts for MN-major tf32 operands
// (pairs with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA side).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

// lo = rn_tf32(x - trunc_tf32(x)): what the tensor core does NOT see of x (it truncates the low 13 mantissa bits),
// itself rounded to nearest tf32 so that the hardware truncation of the lo operand is a no-op:
// |x - hi - lo| <= 2^-22 |x|, unbiased.
__device__ __forceinline__ float lo_part(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(r));
  return __uint_as_float(t);
}

struct Params {
  float* C; const float* D; float* C2; const float* D2;
  float* Clo; float* C2lo;                  // optional "lo" companions of C / C2 (same leading dims) for a following product
  int64_t ldc, ldd, ldc2, ldd2;
  int64_t split_stride;                     // gridDim.z > 1: split z writes its raw partial sum to C + z*split_stride
  int M, N, K;
  float alpha, beta;
  int a_tri, c_lower, b_kmajor, chunk;     // chunk: k-blocks per tensor-core accumulation chain
  // fp16 variant only: operands are x*s split into two halves; the accumulators are multiplied by *ab_inv = 1/(sA*sB)
  // first, and the optional outputs Ch/Cl (C2h/C2l) are the two-half split of C * *c_scale (C2 * *c2_scale)
  const float* ab_inv; const float* c_scale; const float* c2_scale;
  __half* Ch; __half* Cl; __half* C2h; __half* C2l;
  int64_t ldch, ldc2h;
  int nz;                                   // persistent kernels: number of split-K slices (the others use gridDim.z)
  long long* trace; int trace_cap;          // persistent kernels, optional: SM-clock stamps per (pair, item), see set_tc_trace
};

// x -> (hi, lo) fp16 with hi + lo = x to 2^-22 |x| (normal range): the same 22 significand bits as the tf32 split
__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ void store_half4(__half* dst, const float (&x)[4], __half* dst_lo) {
  __half h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_half(x[i], h[i], l[i]);
  *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(dst_lo) = *reinterpret_cast<const uint2*>(l);
}

template <int CG, bool H, int BNT = BN>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap& mapAh, const CUtensorMap& mapAl, const CUtensorMap& mapBh,
                                             const CUtensorMap& mapBl, const Params& p) {
  using G = Geo<CG, BNT>;
  constexpr int STAGES = G::STAGES, STAGE_BYTES = G::STAGE_BYTES, B_BYTES = G::B_BYTES;
  constexpr int BNL = G::BN_LOCAL;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;      // [2]       MMA -> epilogue (chunk finished)
  uint64_t* tempty = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA (buffer drained)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;        // position in the CTA pair; rank 0 issues the MMAs
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BNT;
  const int m0p = CG == 2 ? (m0 - (int)rank * BM) : m0;          // first row of the pair's 256-row tile
  if (p.c_lower && n0 >= m0p + BM * CG) return;                  // tile strictly above the diagonal (pair-uniform)

  // k-block range that can be non-zero for this (pair of) row tile(s)
  constexpr int BKE = H ? BKH : BK;                              // elements per k-block (128 bytes either way)
  constexpr int CB = H ? 64 : 32;                                // columns per MN-major column block (128 bytes)
  const int nkb = (p.K + BKE - 1) / BKE;
  int kb0 = 0, kb1 = nkb;
  if (p.a_tri == 1) kb1 = min(nkb, (m0p + BM * CG + BKE - 1) / BKE);  // lower: k <= row
  if (p.a_tri == 2) kb0 = m0p / BKE;                                  // upper: k >= row
  if (gridDim.z > 1) {                                           // split-K: this CTA takes an even share of the range
    const int tot = max(kb1 - kb0, 0), per = (tot + gridDim.z - 1) / gridDim.z;
    kb0 = kb0 + blockIdx.z * per;
    kb1 = min(kb1, kb0 + per);
  }
  const int nk = max(kb1 - kb0, 0);
  const int nchunks = (nk + p.chunk - 1) / p.chunk;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], G::EPI * CG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if constexpr (BNT == BN) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync();                         // the peer's barriers must exist before anyone signals them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES, kb = kb0 + i;
        if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
        unsigned char* st = smem + s * STAGE_BYTES;
        const int nb = n0 + (int)rank * BNL;                    // this CTA's share of the B columns
        if constexpr (CG == 1) {
          mbar_expect_tx(&full[s], STAGE_BYTES);
          tma_load_2d(st, &mapAh, &full[s], kb * BKE, m0);
          tma_load_2d(st + A_BYTES, &mapAl, &full[s], kb * BKE, m0);
          if (p.b_kmajor) {
            tma_load_2d(st + 2 * A_BYTES, &mapBh, &full[s], kb * BKE, nb);
            tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapBl, &full[s], kb * BKE, nb);
          } else {
            tma_load_3d(st + 2 * A_BYTES, &mapBh, &full[s], 0, kb * BKE, nb / CB);
            tma_load_3d(st + 2 * A_BYTES + B_BYTES, &mapBl, &full[s], 0, kb * BKE, nb / CB);
          }
        } else {
          if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);     // both CTAs' bytes land on the leader's barrier
          tma2_load_2d(st, &mapAh, &full[s], kb * BKE, m0);
          tma2_load_2d(st + A_BYTES, &mapAl, &full[s], kb * BKE, m0);
          if (p.b_kmajor) {
            tma2_load_2d(st + 2 * A_BYTES, &mapBh, &full[s], kb * BKE, nb);
            tma2_load_2d(st + 2 * A_BYTES + B_BYTES, &mapBl, &full[s], kb * BKE, nb);
          } else {
            tma2_load_3d(st + 2 * A_BYTES, &mapBh, &full[s], 0, kb * BKE, nb / CB);
            tma2_load_3d(st + 2 * A_BYTES + B_BYTES, &mapBl, &full[s], 0, kb * BKE, nb / CB);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      // instruction descriptor: D=f32, A=B=tf32, A K-major, B per flag, N=256, M=128 (256 for a CTA pair)
      // (kind::f16: A = B = f16 is format 0)
      const uint32_t fmt = H ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((p.b_kmajor ? 0u : 1u) << 16) | ((uint32_t)(BNT >> 3) << 17) |
                             ((uint32_t)((BM * CG) >> 4) << 24);
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES, c = i / p.chunk, buf = c & 1;
        const bool chunk_start = (i % p.chunk) == 0;
        if (chunk_start && c >= 2) {
          mbar_wait(&tempty[buf], ((c >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mbar_wait(&full[s], (i / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * BNT;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          // A: K-major, rows of 128 B, 8-row groups 1024 B apart; a k-step advances 32 B inside the swizzled row
          const uint64_t ah = umma_desc(st + ks * 32, 16, 1024);
          const uint64_t al = umma_desc(st + A_BYTES + ks * 32, 16, 1024);
          uint64_t bh, bl;
          if (p.b_kmajor) {
            bh = umma_desc(st + 2 * A_BYTES + ks * 32, 16, 1024);
            bl = umma_desc(st + 2 * A_BYTES + B_BYTES + ks * 32, 16, 1024);
          } else if constexpr (H) {
            // B: MN-major fp16, plain 128B swizzle: 64-half column blocks BKH*128 B apart (LBO); a k-step is 16 k-rows of
            // 128 B = two 8-row swizzle atoms 1024 B apart (SBO)
            bh = umma_desc(st + 2 * A_BYTES + ks * 2048, BKH * 128, 1024, 2);
            bl = umma_desc(st + 2 * A_BYTES + B_BYTES + ks * 2048, BKH * 128, 1024, 2);
          } else {
            // B: MN-major (32B-atom swizzle): 32-float column blocks BK*128 B apart (LBO); a k-step is 8 k-rows of
            // 128 B = two 4-row swizzle atoms 512 B apart (SBO)
            bh = umma_desc(st + 2 * A_BYTES + ks * 1024, BK * 128, 512, 1);
            bl = umma_desc(st + 2 * A_BYTES + B_BYTES + ks * 1024, BK * 128, 512, 1);
          }
          const uint32_t first = (chunk_start && ks == 0) ? 0u : 1u;
          if constexpr (CG == 1 && !H) {
            umma_tf32(d_tmem, al, bh, idesc, first);   // small terms first
            umma_tf32(d_tmem, ah, bl, idesc, 1u);
            umma_tf32(d_tmem, ah, bh, idesc, 1u);
          } else if constexpr (CG == 2 && !H) {
            umma2_tf32(d_tmem, al, bh, idesc, first);
            umma2_tf32(d_tmem, ah, bl, idesc, 1u);
            umma2_tf32(d_tmem, ah, bh, idesc, 1u);
          } else if constexpr (CG == 1) {
            umma_f16(d_tmem, al, bh, idesc, first);
            umma_f16(d_tmem, ah, bl, idesc, 1u);
            umma_f16(d_tmem, ah, bh, idesc, 1u);
          } else {
            umma2_f16(d_tmem, al, bh, idesc, first);
            umma2_f16(d_tmem, ah, bl, idesc, 1u);
            umma2_f16(d_tmem, ah, bh, idesc, 1u);
          }
        }
        const bool chunk_end = (i % p.chunk) == p.chunk - 1 || i == nk - 1;
        if constexpr (CG == 1) {
          umma_commit(&empty[s]);                                   // smem stage free once these MMAs retire
          if (chunk_end) umma_commit(&tfull[buf]);                  // chunk finished
        } else {
          umma2_commit_both(&empty[s]);                             // ... in both CTAs of the pair
          if (chunk_end) umma2_commit_both(&tfull[buf]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int h = e >> 2;                   // column half
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&tfull[buf], (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BNT + h * 128);
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        float v[32];
        tmem_ld32(taddr + part * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[part * 32 + i] += v[i];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tempty[buf]);
        else mbar_arrive_cta0(&tempty[buf]);
      }
    }
    // Stage this warp's 32 x 128 block through shared memory (the operand ring is idle: the last chunk is complete, so
    // every TMA load has landed and every MMA has retired) so that global traffic is row-contiguous: one warp
    // instruction = 512 contiguous bytes of one output row.
    constexpr int LDE = 132;
    float* stage = reinterpret_cast<float*>(smem) + e * (32 * LDE);
#pragma unroll
    for (int i = 0; i < 128; i += 4)
      *reinterpret_cast<float4*>(stage + lane * LDE + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    __syncwarp();
    const int col = n0 + h * 128 + lane * 4;
    float* Cb = p.C ? p.C + (int64_t)blockIdx.z * p.split_stride : nullptr;
    const bool raw = gridDim.z > 1;
    auto ok16 = [](const void* q, int64_t ld) { return q == nullptr || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)); };
    const bool al16 = (col + 3 < p.N) && ok16(Cb, p.ldc) && ok16(p.D, p.ldd) && ok16(p.C2, p.ldc2) && ok16(p.D2, p.ldd2) &&
                      ok16(p.Clo, p.ldc) && ok16(p.C2lo, p.ldc2) && ok16(p.Ch, p.ldch) && ok16(p.Cl, p.ldch) &&
                      ok16(p.C2h, p.ldc2h) && ok16(p.C2l, p.ldc2h);
    float inv = 1.f, cs = 1.f, c2s = 1.f;
    if constexpr (H) {
      if (!raw) {
        inv = *p.ab_inv;
        if (p.Ch) cs = *p.c_scale;
        if (p.C2h) c2s = *p.c2_scale;
      }
    }
    // addend rows (D, D2) are fetched 8 rows ahead: one L2/HBM round trip per 8 rows instead of one per row (the
    // epilogue of the products with an addend was latency-bound: +0.5 ms on the C3 B' = E^T A product)
    constexpr int PF = 8;
    const bool need_d = !raw && al16 && p.beta != 0.f, need_e = !raw && al16 && p.D2 != nullptr;
    auto do_row = [&](int r, int row, const float4& d4in, const float4& e4in) {
      const float4 a4 = *reinterpret_cast<const float4*>(stage + r * LDE + lane * 4);
      float v[4] = {a4.x, a4.y, a4.z, a4.w};
      if (raw) {                                  // split-K partial sums: unscaled, summed by splitk_reduce_kernel
        float* crow = Cb + (int64_t)row * p.ldc;
        if (al16) *reinterpret_cast<float4*>(crow + col) = a4;
        else
          for (int i = 0; i < 4; ++i)
            if (col + i < p.N) crow[col + i] = v[i];
        return;
      }
      const float* drow = p.D ? p.D + (int64_t)row * p.ldd : (Cb ? Cb + (int64_t)row * p.ldc : nullptr);
      float o[4], o2[4];
      if (al16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = p.alpha * (H ? v[i] * inv : v[i]);
        if (p.beta != 0.f) {
          const float4 d4 = d4in;
          o[0] += p.beta * d4.x; o[1] += p.beta * d4.y; o[2] += p.beta * d4.z; o[3] += p.beta * d4.w;
        }
        if (Cb) *reinterpret_cast<float4*>(Cb + (int64_t)row * p.ldc + col) = make_float4(o[0], o[1], o[2], o[3]);
        if (p.Clo)
          *reinterpret_cast<float4*>(p.Clo + (int64_t)row * p.ldc + col) = make_float4(lo_part(o[0]), lo_part(o[1]), lo_part(o[2]), lo_part(o[3]));
        if constexpr (H) {
          if (p.Ch) {
            const float xs[4] = {o[0] * cs, o[1] * cs, o[2] * cs, o[3] * cs};
            store_half4(p.Ch + (int64_t)row * p.ldch + col, xs, p.Cl + (int64_t)row * p.ldch + col);
          }
        }
        if (p.D2) {
          const float4 e4 = e4in;
          o2[0] = o[0] + e4.x; o2[1] = o[1] + e4.y; o2[2] = o[2] + e4.z; o2[3] = o[3] + e4.w;
          if (p.C2) *reinterpret_cast<float4*>(p.C2 + (int64_t)row * p.ldc2 + col) = make_float4(o2[0], o2[1], o2[2], o2[3]);
          if (p.C2lo)
            *reinterpret_cast<float4*>(p.C2lo + (int64_t)row * p.ldc2 + col) =
                make_float4(lo_part(o2[0]), lo_part(o2[1]), lo_part(o2[2]), lo_part(o2[3]));
          if constexpr (H) {
            if (p.C2h) {
              const float xs[4] = {o2[0] * c2s, o2[1] * c2s, o2[2] * c2s, o2[3] * c2s};
              store_half4(p.C2h + (int64_t)row * p.ldc2h + col, xs, p.C2l + (int64_t)row * p.ldc2h + col);
            }
          }
        }
      } else {
        for (int i = 0; i < 4; ++i) {
          if (col + i >= p.N) break;
          float oo = p.alpha * (H ? v[i] * inv : v[i]);
          if (p.beta != 0.f) oo += p.beta * drow[col + i];
          if (Cb) Cb[(int64_t)row * p.ldc + col + i] = oo;
          if (p.Clo) p.Clo[(int64_t)row * p.ldc + col + i] = lo_part(oo);
          if constexpr (H) {
            if (p.Ch) split_half(oo * cs, p.Ch[(int64_t)row * p.ldch + col + i], p.Cl[(int64_t)row * p.ldch + col + i]);
          }
          if (p.D2) {
            const float oo2 = oo + p.D2[(int64_t)row * p.ldd2 + col + i];
            if (p.C2) p.C2[(int64_t)row * p.ldc2 + col + i] = oo2;
            if (p.C2lo) p.C2lo[(int64_t)row * p.ldc2 + col + i] = lo_part(oo2);
            if constexpr (H) {
              if (p.C2h) split_half(oo2 * c2s, p.C2h[(int64_t)row * p.ldc2h + col + i], p.C2l[(int64_t)row * p.ldc2h + col + i]);
            }
          }
        }
      }
    };
    if (!need_d && !need_e) {                     // nothing to fetch: the plain row loop (less code on the hot products)
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int r = 0; r < 32; ++r) {
        const int row = m0 + q * 32 + r;
        if (row >= p.M) break;
        do_row(r, row, z4, z4);
      }
    } else
    for (int r0 = 0; r0 < 32; r0 += PF) {
      const int row0 = m0 + q * 32 + r0;
      if (row0 >= p.M) break;
      float4 dpre[PF], epre[PF];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        dpre[u] = epre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + u < p.M) {
          if (need_d)
            dpre[u] = *reinterpret_cast<const float4*>((p.D ? p.D + (int64_t)(row0 + u) * p.ldd : Cb + (int64_t)(row0 + u) * p.ldc) + col);
          if (need_e) epre[u] = *reinterpret_cast<const float4*>(p.D2 + (int64_t)(row0 + u) * p.ldd2 + col);
        }
      }
#pragma unroll
      for (int u = 0; u < PF; ++u)
        if (row0 + u < p.M) do_row(r0 + u, row0 + u, dpre[u], epre[u]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 2) cluster_sync();                         // the peer may still be reading operands / TMEM
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else if constexpr (BNT == BN) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p) {
  gemm_tc_body<1, false>(mapAh, mapAl, mapBh, mapBl, p);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p) {
  gemm_tc_body<2, false>(mapAh, mapAl, mapBh, mapBl, p);
}

// 3xFP16 variants: the same pipeline on tcgen05 kind::f16 (twice the K per 128-byte smem row and per MMA instruction)
__global__ void __launch_bounds__(THREADS, 1)
gemm_tch_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p) {
  gemm_tc_body<1, true>(mapAh, mapAl, mapBh, mapBl, p);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tch2_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                 const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p) {
  gemm_tc_body<2, true>(mapAh, mapAl, mapBh, mapBl, p);
}

// 3xFP16, CTA pairs, 256 x 128 tiles, two pairs resident per SM pair
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Geo<2, 128>::NTHREADS, 2)
gemm_tch2n_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                  const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p) {
  gemm_tc_body<2, true, 128>(mapAh, mapAl, mapBh, mapBl, p);
}

// =====================================================================================================================
// Persistent CTA-pair variant.  One CTA pair per SM pair lives for the whole product and walks a list of work items
// (256 x 256 output tile, split-K slice) that the host balanced over the pairs (cost = k-blocks of the item, heavy row tiles
// of a column block first so that the column block's operand stays in L2 while its row tiles run side by side).  What this
// buys over one CTA pair per tile: barrier / tensor-memory set-up once per SM instead of once per tile, the producer warp
// keeps the operand ring full ACROSS tiles (no pipeline fill per tile), and the MMA warp starts the next tile's first two
// accumulation chunks while the epilogue warps are still storing the previous tile -- the store phase goes through a
// per-warp 4 KB transposition strip that does NOT alias the operand ring (3 stages of 64 KB + 8 x 4 KB + barriers = 225 KB).
// 384 threads = three warpgroups, so that registers can be moved between the roles (setmaxnreg works on warpgroups): warps 0-3
// (TMA producer, MMA issuer, two idle) shrink to 40 registers and the eight epilogue warps 4-11 grow to 232 -- the 128 master sums
// per epilogue thread do not fit the 168 registers a 10- or 12-warp CTA gets at launch, and a spilled master sum costs an L2 round
// trip per accumulation chunk (the shared-memory carve-out leaves no L1).
constexpr int P_THREADS = 384;
constexpr int P_STAGES = 3;
constexpr int P_STRIP_FLOATS = 32 * 32;                                  // per epilogue warp: 32 rows x 32 columns, XOR-swizzled
constexpr int P_SMEM_BYTES = P_STAGES * Geo<2>::STAGE_BYTES + EPI_WARPS * P_STRIP_FLOATS * 4 + 1024 /*align*/ + 256 /*barriers*/;
static_assert(P_SMEM_BYTES <= 227 * 1024, "persistent GEMM: shared memory");

// a work item: {pair-tile row, tile column, split-K slice, k-block range kb0 | kb1 << 16 (kb1 == kb0: write zeros)}
__device__ __forceinline__ void item_krange(const int4& item, int& kb0, int& kb1) {
  kb0 = item.w & 0xFFFF;
  kb1 = max((int)((uint32_t)item.w >> 16), kb0);
}

struct EpiConst {
  float* Cb;            // C (+ split-K slice offset)
  bool raw, vec;        // raw split-K partial sums; every pointer / leading dimension allows 16-byte accesses
  float inv, cs, c2s;
};

// one 4-column piece of an output row: scale, addends, every requested output (fp32, lo, two-half splits)
template <bool H>
__device__ __forceinline__ void emit4(const Params& p, const EpiConst& ec, int row, int col, const float4& a4, const float4& d4,
                                      const float4& e4) {
  const float v[4] = {a4.x, a4.y, a4.z, a4.w};
  const bool al16 = ec.vec && (col + 3 < p.N);
  if (ec.raw) {                                   // split-K partial sums: unscaled, summed by splitk_reduce_kernel
    float* crow = ec.Cb + (int64_t)row * p.ldc;
    if (al16) *reinterpret_cast<float4*>(crow + col) = a4;
    else
      for (int i = 0; i < 4; ++i)
        if (col + i < p.N) crow[col + i] = v[i];
    return;
  }
  float* Cb = ec.Cb;
  if (al16) {
    float o[4], o2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = p.alpha * (H ? v[i] * ec.inv : v[i]);
    if (p.beta != 0.f) { o[0] += p.beta * d4.x; o[1] += p.beta * d4.y; o[2] += p.beta * d4.z; o[3] += p.beta * d4.w; }
    if (Cb) *reinterpret_cast<float4*>(Cb + (int64_t)row * p.ldc + col) = make_float4(o[0], o[1], o[2], o[3]);
    if (p.Clo)
      *reinterpret_cast<float4*>(p.Clo + (int64_t)row * p.ldc + col) = make_float4(lo_part(o[0]), lo_part(o[1]), lo_part(o[2]), lo_part(o[3]));
    if constexpr (H) {
      if (p.Ch) {
        const float xs[4] = {o[0] * ec.cs, o[1] * ec.cs, o[2] * ec.cs, o[3] * ec.cs};
        store_half4(p.Ch + (int64_t)row * p.ldch + col, xs, p.Cl + (int64_t)row * p.ldch + col);
      }
    }
    if (p.D2) {
      o2[0] = o[0] + e4.x; o2[1] = o[1] + e4.y; o2[2] = o[2] + e4.z; o2[3] = o[3] + e4.w;
      if (p.C2) *reinterpret_cast<float4*>(p.C2 + (int64_t)row * p.ldc2 + col) = make_float4(o2[0], o2[1], o2[2], o2[3]);
      if (p.C2lo)
        *reinterpret_cast<float4*>(p.C2lo + (int64_t)row * p.ldc2 + col) =
            make_float4(lo_part(o2[0]), lo_part(o2[1]), lo_part(o2[2]), lo_part(o2[3]));
      if constexpr (H) {
        if (p.C2h) {
          const float xs[4] = {o2[0] * ec.c2s, o2[1] * ec.c2s, o2[2] * ec.c2s, o2[3] * ec.c2s};
          store_half4(p.C2h + (int64_t)row * p.ldc2h + col, xs, p.C2l + (int64_t)row * p.ldc2h + col);
        }
      }
    }
  } else {
    const float* drow = p.D ? p.D + (int64_t)row * p.ldd : (Cb ? Cb + (int64_t)row * p.ldc : nullptr);
    for (int i = 0; i < 4; ++i) {
      if (col + i >= p.N) break;
      float oo = p.alpha * (H ? v[i] * ec.inv : v[i]);
      if (p.beta != 0.f) oo += p.beta * drow[col + i];
      if (Cb) Cb[(int64_t)row * p.ldc + col + i] = oo;
      if (p.Clo) p.Clo[(int64_t)row * p.ldc + col + i] = lo_part(oo);
      if constexpr (H) {
        if (p.Ch) split_half(oo * ec.cs, p.Ch[(int64_t)row * p.ldch + col + i], p.Cl[(int64_t)row * p.ldch + col + i]);
      }
      if (p.D2) {
        const float oo2 = oo + p.D2[(int64_t)row * p.ldd2 + col + i];
        if (p.C2) p.C2[(int64_t)row * p.ldc2 + col + i] = oo2;
        if (p.C2lo) p.C2lo[(int64_t)row * p.ldc2 + col + i] = lo_part(oo2);
        if constexpr (H) {
          if (p.C2h) split_half(oo2 * ec.c2s, p.C2h[(int64_t)row * p.ldc2h + col + i], p.C2l[(int64_t)row * p.ldc2h + col + i]);
        }
      }
    }
  }
}

template <bool H>
__device__ __forceinline__ void gemm_tcp_body(const CUtensorMap& mapAh, const CUtensorMap& mapAl, const CUtensorMap& mapBh,
                                              const CUtensorMap& mapBl, const Params& p, const int* __restrict__ offs,
                                              const int4* __restrict__ items) {
  using G = Geo<2>;
  constexpr int STAGES = P_STAGES, STAGE_BYTES = G::STAGE_BYTES, B_BYTES = G::B_BYTES, BNL = G::BN_LOCAL;
  constexpr int BKE = H ? BKH : BK;                              // elements per k-block (128 bytes either way)
  constexpr int CB = H ? 64 : 32;                                // columns per MN-major column block (128 bytes)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* strips = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_WARPS * P_STRIP_FLOATS * 4);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * STAGES;      // [2]       MMA -> epilogue (chunk finished)
  uint64_t* tempty = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA (buffer drained)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                       // position in the CTA pair; rank 0 issues the MMAs
  const int pair = blockIdx.x >> 1;
  const int w0 = offs[pair], w1 = offs[pair + 1];                // this pair's work items

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], EPI_WARPS * 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();                                                // the peer's barriers must exist before anyone signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // (each setmaxnreg sits at the top of the branch it governs: after a join ptxas assumes the smaller of the two limits)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (runs ahead across tiles)
    if (lane == 0) {
      uint32_t it = 0;                                           // k-blocks loaded so far (ring position)
      for (int w = w0; w < w1; ++w) {
        const int4 item = __ldg(items + w);
        const int m0p = item.x * (2 * BM), n0 = item.y * BN;
        int kb0, kb1;
        item_krange(item, kb0, kb1);
        const int m0 = m0p + (int)rank * BM, nb = n0 + (int)rank * BNL;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES, round = it / STAGES;
          if (round >= 1) mbar_wait(&empty[s], (round - 1) & 1);
          unsigned char* st = smem + s * STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);     // both CTAs' bytes land on the leader's barrier
          tma2_load_2d(st, &mapAh, &full[s], kb * BKE, m0);
          tma2_load_2d(st + A_BYTES, &mapAl, &full[s], kb * BKE, m0);
          if (p.b_kmajor) {
            tma2_load_2d(st + 2 * A_BYTES, &mapBh, &full[s], kb * BKE, nb);
            tma2_load_2d(st + 2 * A_BYTES + B_BYTES, &mapBl, &full[s], kb * BKE, nb);
          } else {
            tma2_load_3d(st + 2 * A_BYTES, &mapBh, &full[s], 0, kb * BKE, nb / CB);
            tma2_load_3d(st + 2 * A_BYTES + B_BYTES, &mapBl, &full[s], 0, kb * BKE, nb / CB);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      const uint32_t fmt = H ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((p.b_kmajor ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((BM * 2) >> 4) << 24);
      uint32_t it = 0, gc = 0;                                   // k-blocks / accumulation chunks issued so far
      for (int w = w0; w < w1; ++w) {
        const int4 item = __ldg(items + w);
        int kb0, kb1;
        item_krange(item, kb0, kb1);
        const int nk = kb1 - kb0;
        const bool tr = p.trace != nullptr && w < p.trace_cap;
        if (tr) p.trace[(size_t)w * 8 + 0] = clock64();                    // MMA warp reaches the item
        for (int i = 0; i < nk; ++i, ++it) {
          const uint32_t s = it % STAGES, c = gc + (uint32_t)(i / p.chunk), buf = c & 1;
          const bool chunk_start = (i % p.chunk) == 0;
          if (chunk_start && c >= 2) {
            mbar_wait(&tempty[buf], ((c >> 1) - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          mbar_wait(&full[s], (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t d_tmem = tmem_base + buf * BN;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t ah = umma_desc(st + ks * 32, 16, 1024);
            const uint64_t al = umma_desc(st + A_BYTES + ks * 32, 16, 1024);
            uint64_t bh, bl;
            if (p.b_kmajor) {
              bh = umma_desc(st + 2 * A_BYTES + ks * 32, 16, 1024);
              bl = umma_desc(st + 2 * A_BYTES + B_BYTES + ks * 32, 16, 1024);
            } else if constexpr (H) {
              bh = umma_desc(st + 2 * A_BYTES + ks * 2048, BKH * 128, 1024, 2);
              bl = umma_desc(st + 2 * A_BYTES + B_BYTES + ks * 2048, BKH * 128, 1024, 2);
            } else {
              bh = umma_desc(st + 2 * A_BYTES + ks * 1024, BK * 128, 512, 1);
              bl = umma_desc(st + 2 * A_BYTES + B_BYTES + ks * 1024, BK * 128, 512, 1);
            }
            const uint32_t first = (chunk_start && ks == 0) ? 0u : 1u;
            if constexpr (H) {
              umma2_f16(d_tmem, al, bh, idesc, first);         // small terms first
              umma2_f16(d_tmem, ah, bl, idesc, 1u);
              umma2_f16(d_tmem, ah, bh, idesc, 1u);
            } else {
              umma2_tf32(d_tmem, al, bh, idesc, first);
              umma2_tf32(d_tmem, ah, bl, idesc, 1u);
              umma2_tf32(d_tmem, ah, bh, idesc, 1u);
            }
          }
          const bool chunk_end = (i % p.chunk) == p.chunk - 1 || i == nk - 1;
          umma2_commit_both(&empty[s]);                          // smem stage free (in both CTAs) once these MMAs retire
          if (chunk_end) umma2_commit_both(&tfull[buf]);         // chunk finished
          if (tr && i == 0) p.trace[(size_t)w * 8 + 1] = clock64();          // first k-block issued (operands had landed)
        }
        if (tr) p.trace[(size_t)w * 8 + 2] = clock64();                    // last k-block issued
        gc += (uint32_t)((nk + p.chunk - 1) / p.chunk);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 4;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int h = e >> 2;                   // column half
    uint32_t gc = 0;
    for (int w = w0; w < w1; ++w) {
      int nchunks;
      {
        const int4 item = __ldg(items + w);
        int kb0, kb1;
        item_krange(item, kb0, kb1);
        nchunks = (kb1 - kb0 + p.chunk - 1) / p.chunk;
      }
      float acc[128];
#pragma unroll
      for (int i = 0; i < 128; ++i) acc[i] = 0.f;
      const bool tr = p.trace != nullptr && w < p.trace_cap && e == 0 && lane == 0 && rank == 0;
      if (tr) p.trace[(size_t)w * 8 + 3] = clock64();                      // epilogue warp reaches the item
      for (int cc = 0; cc < nchunks; ++cc) {
        const uint32_t c = gc + (uint32_t)cc, buf = c & 1;
        mbar_wait(&tfull[buf], (c >> 1) & 1);
        if (tr && cc == 0) p.trace[(size_t)w * 8 + 4] = clock64();           // first chunk complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * 128);
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          float v[32];
          tmem_ld32(taddr + part * 32, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[part * 32 + i] += v[i];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cta0(&tempty[buf]);
      }
      gc += (uint32_t)nchunks;
      if (tr) p.trace[(size_t)w * 8 + 5] = clock64();                      // last chunk added: store phase starts
      // ---- store phase: both TMEM buffers are already released, so the MMA warp is on the next item's first chunks.
      // Per 32-column part: lane r writes row r of the part into the strip (16-byte pieces XOR-swizzled by row: conflict-free),
      // then 8 lanes x 16 bytes read one row back, so a warp store instruction covers 4 rows x 128 contiguous bytes.
      // (everything the store phase needs is derived here, not before the chunk loop: the master sums leave few registers)
      const int4 item = __ldg(items + w);
      const int m0p = item.x * (2 * BM), n0 = item.y * BN;
      float* strip = strips + e * P_STRIP_FLOATS;
      EpiConst ec;
      ec.raw = p.nz > 1;
      auto ok16 = [](const void* q_, int64_t ld) { return q_ == nullptr || (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(q_) & 15) == 0)); };
      ec.vec = ok16(p.C, p.ldc) && ok16(p.D, p.ldd) && ok16(p.C2, p.ldc2) && ok16(p.D2, p.ldd2) && ok16(p.Clo, p.ldc) &&
               ok16(p.C2lo, p.ldc2) && ok16(p.Ch, p.ldch) && ok16(p.Cl, p.ldch) && ok16(p.C2h, p.ldc2h) && ok16(p.C2l, p.ldc2h) &&
               ((p.split_stride & 3) == 0);
      ec.inv = ec.cs = ec.c2s = 1.f;
      if constexpr (H) {
        if (!ec.raw) {
          ec.inv = *p.ab_inv;
          if (p.Ch) ec.cs = *p.c_scale;
          if (p.C2h) ec.c2s = *p.c2_scale;
        }
      }
      const bool need_d = !ec.raw && ec.vec && p.beta != 0.f, need_e = !ec.raw && ec.vec && p.D2 != nullptr;
      ec.Cb = p.C ? p.C + (int64_t)item.z * p.split_stride : nullptr;
      const int rbase = m0p + (int)rank * BM + q * 32;
      const int jj = lane & 7, rsub = lane >> 3;
      auto put = [&](auto part_c) {                      // acc is register-resident: the part index must be a constant
        constexpr int PT = decltype(part_c)::value;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(strip + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_float4(acc[PT * 32 + 4 * j], acc[PT * 32 + 4 * j + 1], acc[PT * 32 + 4 * j + 2], acc[PT * 32 + 4 * j + 3]);
      };
      const bool addends = need_d || need_e;
      const bool fast = ec.vec && !addends && p.beta == 0.f && p.Clo == nullptr && p.D2 == nullptr && rbase + 32 <= p.M &&
                        n0 + h * 128 + 128 <= p.N;
#pragma unroll 1
      for (int part = 0; part < 4; ++part) {             // rolled: ONE copy of the store code (it runs once per tile)
        switch (part) {
          case 0: put(std::integral_constant<int, 0>{}); break;
          case 1: put(std::integral_constant<int, 1>{}); break;
          case 2: put(std::integral_constant<int, 2>{}); break;
          default: put(std::integral_constant<int, 3>{}); break;
        }
        __syncwarp();
        const int col = n0 + h * 128 + part * 32 + jj * 4;
        if (fast) {
          // interior tile, no addends: the 8 store instructions of the part are independent -- all 8 strip reads first, then the
          // arithmetic, then the stores, so the latencies overlap (a rolled row loop is latency-bound with 2 warps per scheduler:
          // measured 9.5 us per tile for 3 output streams, and it is exposed because the MMA warp is only two chunks ahead)
          float4 a[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int r = t * 4 + rsub;
            a[t] = *reinterpret_cast<const float4*>(strip + r * 32 + ((jj ^ (r & 7)) << 2));
          }
          if (!ec.raw) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              a[t].x = p.alpha * (H ? a[t].x * ec.inv : a[t].x);
              a[t].y = p.alpha * (H ? a[t].y * ec.inv : a[t].y);
              a[t].z = p.alpha * (H ? a[t].z * ec.inv : a[t].z);
              a[t].w = p.alpha * (H ? a[t].w * ec.inv : a[t].w);
            }
          }
          const int64_t row0 = rbase + rsub;
          if (ec.Cb) {
            float* c = ec.Cb + row0 * p.ldc + col;
#pragma unroll
            for (int t = 0; t < 8; ++t) *reinterpret_cast<float4*>(c + (int64_t)(4 * t) * p.ldc) = a[t];
          }
          if constexpr (H) {
            if (p.Ch && !ec.raw) {
              __half* ch = p.Ch + row0 * p.ldch + col;
              __half* cl = p.Cl + row0 * p.ldch + col;
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float x0 = a[t].x * ec.cs, x1 = a[t].y * ec.cs, x2 = a[t].z * ec.cs, x3 = a[t].w * ec.cs;
                const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(ch + (int64_t)(4 * t) * p.ldch) = hv;
                *reinterpret_cast<uint2*>(cl + (int64_t)(4 * t) * p.ldch) = lv;
              }
            }
          }
        } else {
          // general path (edge tiles, addends, lo companions): one row group per iteration; addend rows (D, D2) are fetched 2 store
          // instructions (8 rows) ahead -- the epilogue of the products with an addend is otherwise bound by one L2 / HBM round trip per row
          auto fetch = [&](int t, float4& d, float4& ee) {
            const int row = rbase + t * 4 + rsub;
            d = ee = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < 8 && row < p.M && col + 3 < p.N) {
              if (need_d) d = *reinterpret_cast<const float4*>((p.D ? p.D + (int64_t)row * p.ldd : ec.Cb + (int64_t)row * p.ldc) + col);
              if (need_e) ee = *reinterpret_cast<const float4*>(p.D2 + (int64_t)row * p.ldd2 + col);
            }
          };
          float4 d0, d1, e0, e1;
          d0 = d1 = e0 = e1 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (addends) {
            fetch(0, d0, e0);
            fetch(1, d1, e1);
          }
#pragma unroll 1
          for (int t = 0; t < 8; ++t) {
            const float4 dc = d0, ecur = e0;
            d0 = d1, e0 = e1;
            if (addends) fetch(t + 2, d1, e1);
            const int r = t * 4 + rsub, row = rbase + r;
            const float4 a4 = *reinterpret_cast<const float4*>(strip + r * 32 + ((jj ^ (r & 7)) << 2));
            if (row < p.M && col < p.N) emit4<H>(p, ec, row, col, a4, dc, ecur);
          }
        }
        __syncwarp();
      }
      if (tr) p.trace[(size_t)w * 8 + 6] = clock64();                      // store phase done
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();                                                // the peer may still be reading operands / TMEM
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_tc2p_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                 const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p,
                 const int* __restrict__ offs, const int4* __restrict__ items) {
  gemm_tcp_body<false>(mapAh, mapAl, mapBh, mapBl, p, offs, items);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_tch2p_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                  const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const Params p,
                  const int* __restrict__ offs, const int4* __restrict__ items) {
  gemm_tcp_body<true>(mapAh, mapAl, mapBh, mapBl, p, offs, items);
}

__global__ void split_lo_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ lo, int64_t ldl, int rows, int cols) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4, i = blockIdx.y;
  if (i >= rows || j >= cols) return;
  const float* xr = x + (int64_t)i * ldx;
  float* lr = lo + (int64_t)i * ldl;
  if (j + 3 < cols && ((ldx | ldl) & 3) == 0) {
    const float4 v = *reinterpret_cast<const float4*>(xr + j);
    float4 o;
    o.x = lo_part(v.x);
    o.y = lo_part(v.y);
    o.z = lo_part(v.z);
    o.w = lo_part(v.w);
    *reinterpret_cast<float4*>(lr + j) = o;
  } else {
    for (int q = j; q < min(j + 4, cols); ++q) lr[q] = lo_part(xr[q]);
  }
}

// C = alpha * sum_z part[z] + beta * D, summed in fp64; lower != 0: only col <= row is touched
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int nsplit, int64_t stride, int64_t ldp, int M, int N,
                                     float alpha, float beta, const float* __restrict__ D, int64_t ldd, float* __restrict__ C,
                                     int64_t ldc, int lower, const float* __restrict__ inv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (i >= M || j >= N || (lower && j > i)) return;
  double s = 0.0;
  for (int z = 0; z < nsplit; ++z) s += (double)part[(int64_t)z * stride + (int64_t)i * ldp + j];
  if (inv) s *= (double)*inv;
  float o = alpha * (float)s;
  if (beta != 0.f) o += beta * (D ? D[(int64_t)i * ldd + j] : C[(int64_t)i * ldc + j]);
  C[(int64_t)i * ldc + j] = o;
}

__global__ void transpose_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows, int cols) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bi + r, j = bj + tx;
    tile[r][tx] = (i < rows && j < cols) ? src[(int64_t)i * lds + j] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bj + r, j = bi + tx;       // dst is cols x rows
    if (i < cols && j < rows) dst[(int64_t)i * ldd + j] = tile[tx][r];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// row-major [rows][cols] fp32 (cols contiguous): box = 32 floats x box_rows rows, 128B swizzle
static bool map_kmajor(CUtensorMap* m, const float* base, int64_t ld, int rows, int cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// row-major [K][N] fp32 (N contiguous) viewed as [N/32][K][32]: box = 32 x BK x BN/32 lands as BN/32 column blocks
static bool map_mnmajor(CUtensorMap* m, const float* base, int64_t ld, int K, int N, int box_cols) {
  cuuint64_t gdim[3] = {32, (cuuint64_t)K, (cuuint64_t)((N + 31) / 32)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)(box_cols / 32)};
  cuuint32_t es[3] = {1, 1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp16 versions: box = 64 halves (128 B) x box_rows; K x N row-major viewed as [N/64][K][64], plain 128B swizzle
static bool map_kmajor_h(CUtensorMap* m, const __half* base, int64_t ld, int rows, int cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool map_mnmajor_h(CUtensorMap* m, const __half* base, int64_t ld, int K, int N, int box_cols) {
  cuuint64_t gdim[3] = {64, (cuuint64_t)K, (cuuint64_t)((N + 63) / 64)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)BKH, (cuuint32_t)(box_cols / 64)};
  cuuint32_t es[3] = {1, 1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------------------------
// Work lists of the persistent kernels.  Built on the host once per (device, shape, triangle mode, split), kept on the
// device: [offsets of the P pairs | P + 1 ints, padded to 4] [items: int4 {pair-tile row, tile column, split slice, -}].
static long long* g_tc_trace = nullptr;
static int g_tc_trace_cap = 0;
void set_tc_trace(long long* buf, int cap_items) { g_tc_trace = buf; g_tc_trace_cap = buf ? cap_items : 0; }
static int g_tc_max_pairs = 0;     // 0: every CTA pair the device can hold; > 0: at most that many (a product that shares the GPU)
void set_tc_max_pairs(int n) { g_tc_max_pairs = n > 0 ? n : 0; }
int get_tc_max_pairs() { return g_tc_max_pairs; }
static int g_tc_persistent = 1;
void set_tc_persistent(int on) { g_tc_persistent = on ? 1 : 0; }
int get_tc_persistent() { return g_tc_persistent; }

namespace tc {

struct SchedKey {
  int dev, M, N, K, a_tri, c_lower, nz, P, bke;
  bool operator<(const SchedKey& o) const {
    return std::tie(dev, M, N, K, a_tri, c_lower, nz, P, bke) < std::tie(o.dev, o.M, o.N, o.K, o.a_tri, o.c_lower, o.nz, o.P, o.bke);
  }
};
struct SchedEntry { int* dev_ptr = nullptr; int n_off = 0, nz_eff = 1; std::vector<int> host; };

// pure function (also exported for the CPU tests): the table for P pairs.  *nz_eff = split-K slices the reduction has to sum.
//   nz == 1: whole tiles, dealt to the least-loaded pair column block by column block (heavy row tiles first; the last 3P
//            items longest-first so that the pairs finish together).
//   nz > 1 : "stream-K": the k-blocks of all tiles are laid end to end and cut into P equal runs, so every pair gets the same
//            number of k-blocks whatever the tile count; a tile cut by a run boundary has its pieces in consecutive slices,
//            and slices a tile does not use are written as zeros by empty items.  Falls back to nz uniform slices per tile
//            when a tile would need more than nz pieces.
std::vector<int> build_sched(int M, int N, int K, int a_tri, int c_lower, int nz, int P, int bke, int* nz_eff) {
  const int T = ceil_div(M, 2 * BM), NT = ceil_div(N, BN), nkb = ceil_div(K, bke);
  auto range_of = [&](int mt, int& kb0, int& kb1) {
    kb0 = 0, kb1 = nkb;
    if (a_tri == 1) kb1 = std::min(nkb, (mt * 2 * BM + 2 * BM + bke - 1) / bke);
    if (a_tri == 2) kb0 = mt * 2 * BM / bke;
    kb1 = std::max(kb1, kb0);
  };
  struct It { int mt, nt, z, kb0, kb1; };
  constexpr int FIXED = 3;                                          // per-item overhead in k-block units (epilogue hand-over)
  std::vector<std::vector<It>> per(P);
  std::vector<It> tiles;                                            // whole tiles in locality order
  for (int nt = 0; nt < NT; ++nt)                                   // column block by column block (its operand stays in L2) ...
    for (int a = 0; a < T; ++a) {
      const int mt = a_tri == 1 ? T - 1 - a : a;                    // ... heavy row tiles first
      if (c_lower && nt > mt) continue;                             // tile strictly above the diagonal
      It it{mt, nt, 0, 0, 0};
      range_of(mt, it.kb0, it.kb1);
      tiles.push_back(it);
    }
  int slices = 1;
  bool done = false;
  if (nz > 1) {
    long total = 0;
    for (const It& t : tiles) total += t.kb1 - t.kb0;
    std::vector<std::vector<It>> cand(P);
    long pos = 0;                                                   // k-blocks laid out so far
    int q = 0, worst = 1;
    bool ok = total > 0;
    for (const It& t : tiles) {
      int k = t.kb0, z = 0;
      while (k < t.kb1) {
        while (q < P - 1 && pos >= (total * (q + 1)) / P) ++q;      // run of pair q: [total q / P, total (q + 1) / P)
        const long room = (q == P - 1 ? total : (total * (q + 1)) / P) - pos;
        const int take = (int)std::min<long>(t.kb1 - k, std::max<long>(room, 1));
        cand[q].push_back({t.mt, t.nt, z++, k, k + take});
        k += take;
        pos += take;
      }
      if (z == 0) cand[q].push_back({t.mt, t.nt, z++, t.kb0, t.kb0});
      worst = std::max(worst, z);
      if (z > nz) ok = false;
      // (empty items for the slices this tile does not use are appended below, once `worst` is known)
    }
    if (ok) {
      std::vector<int> used((size_t)T * NT, 0);
      std::vector<int> owner((size_t)T * NT, 0);
      for (int qq = 0; qq < P; ++qq)
        for (const It& it : cand[qq]) {
          used[(size_t)it.mt * NT + it.nt] = std::max(used[(size_t)it.mt * NT + it.nt], it.z + 1);
          owner[(size_t)it.mt * NT + it.nt] = qq;
        }
      for (const It& t : tiles)
        for (int z = used[(size_t)t.mt * NT + t.nt]; z < worst; ++z)
          cand[owner[(size_t)t.mt * NT + t.nt]].push_back({t.mt, t.nt, z, t.kb0, t.kb0});
      // every pair walks k upwards (the head of its second tile before the tail of its first): all pairs are then at about the
      // same k at the same time and share the operand rows of that k window through L2 -- with the runs in list order each pair
      // sits at its own k offset and the operands are re-read from HBM for every tile (measured 1.43 instead of 1.15 ms at C3)
      for (auto& v : cand) std::stable_sort(v.begin(), v.end(), [](const It& x, const It& y) { return x.kb0 < y.kb0; });
      per.swap(cand);
      slices = worst;
      done = true;
    }
  }
  if (!done) {
    std::vector<It> list;
    list.reserve(tiles.size() * (size_t)nz);
    for (const It& t : tiles)
      for (int z = 0; z < nz; ++z) {
        It it = t;
        it.z = z;
        if (nz > 1) {
          const int tot = t.kb1 - t.kb0, each = (tot + nz - 1) / nz;
          it.kb0 = std::min(t.kb0 + z * each, t.kb1);
          it.kb1 = std::min(t.kb1, it.kb0 + each);
        }
        list.push_back(it);
      }
    const size_t tail = std::min(list.size(), (size_t)3 * P);       // the last items: longest first, so the pairs finish together
    std::stable_sort(list.end() - tail, list.end(), [](const It& x, const It& y) { return x.kb1 - x.kb0 > y.kb1 - y.kb0; });
    std::vector<long> load(P, 0);
    for (const It& it : list) {
      int best = 0;
      for (int qq = 1; qq < P; ++qq)
        if (load[qq] < load[best]) best = qq;
      per[best].push_back(it);
      load[best] += it.kb1 - it.kb0 + FIXED;
    }
    slices = nz;
  }
  if (nz_eff) *nz_eff = slices;
  size_t count = 0;
  for (const auto& v : per) count += v.size();
  const int n_off = (P + 1 + 3) & ~3;
  std::vector<int> out((size_t)n_off + 4 * count, 0);
  int pos = 0;
  for (int qq = 0; qq < P; ++qq) {
    out[qq] = pos;
    for (const It& it : per[qq]) {
      int* d = &out[(size_t)n_off + 4 * (size_t)pos++];
      d[0] = it.mt; d[1] = it.nt; d[2] = it.z; d[3] = (int)((uint32_t)it.kb0 | ((uint32_t)it.kb1 << 16));
    }
  }
  out[P] = pos;
  return out;
}

static std::map<SchedKey, SchedEntry> g_sched;
static std::mutex g_sched_mu;

// returns false when the table does not exist yet and cannot be created now (stream capture in progress)
static bool get_sched(const SchedKey& k, cudaStream_t st, const int** offs, const int4** items, int* nz_eff) {
  std::lock_guard<std::mutex> lk(g_sched_mu);
  auto itf = g_sched.find(k);
  if (itf == g_sched.end()) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
    SchedEntry e;
    if (ceil_div(k.K, k.bke) > 0xFFFF) return false;              // the packed k-block range of an item is 16 + 16 bits
    e.host = build_sched(k.M, k.N, k.K, k.a_tri, k.c_lower, k.nz, k.P, k.bke, &e.nz_eff);
    e.n_off = (k.P + 1 + 3) & ~3;
    if (cudaMalloc(&e.dev_ptr, e.host.size() * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return false; }
    // synchronous copy: the table is tiny and built once per shape
    if (cudaMemcpy(e.dev_ptr, e.host.data(), e.host.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(e.dev_ptr);
      return false;
    }
    itf = g_sched.emplace(k, std::move(e)).first;
  }
  *offs = itf->second.dev_ptr;
  *nz_eff = itf->second.nz_eff;
  *items = reinterpret_cast<const int4*>(itf->second.dev_ptr + itf->second.n_off);
  return true;
}

// CTA pairs that can be resident at once (one per SM pair unless the part has unpaired SMs)
template <typename Kern>
static int resident_pairs(Kern kern, int dev) {
  static int cache[64] = {};
  if (cache[dev & 63] > 0) return cache[dev & 63];
  int sms = 0, nc = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sms & ~1), 1, 1);
  cfg.blockDim = dim3(P_THREADS, 1, 1);
  cfg.dynamicSmemBytes = P_SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = sms / 2; }
  nc = std::max(1, std::min(nc, sms / 2));
  cache[dev & 63] = nc;
  return nc;
}

}  // namespace tc

static int g_tc_tile_n = 256;    // 3xFP16 CTA-pair kernel: 256 (one pair per SM pair) or 128 (two pairs per SM pair)
void set_tc_tile_n(int n) { g_tc_tile_n = (n == 128) ? 128 : 256; }
int get_tc_tile_n() { return g_tc_tile_n; }
static int g_tc_cta_group = 2;   // CTA pairs by default: +9% over single-CTA tiles on the C3 whitening product
void set_tc_cta_group(int cg) { g_tc_cta_group = (cg == 2) ? 2 : 1; }
int get_tc_cta_group() { return g_tc_cta_group; }

int gemm_tc_supported(const float* A, int64_t lda, const float* B, int64_t ldb, int b_kmajor, int N) {
  if (!tc::encode_fn()) return 0;
  if ((lda & 3) || (ldb & 3)) return 0;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return 0;
  if (!b_kmajor && ldb < (int64_t)((N + 31) / 32) * 32) return 0;     // the last 32-column block must stay inside the row
  return 1;
}

int gemm_tc(const float* Ah, const float* Al, int64_t lda, const float* Bh, const float* Bl, int64_t ldb, int b_kmajor,
            int M, int N, int K, float alpha, float beta, float* C, int64_t ldc, const float* D, int64_t ldd, float* C2,
            int64_t ldc2, const float* D2, int64_t ldd2, int a_tri, int c_lower, int chunk, float* Clo, float* C2lo,
            int nsplit, float* split_ws, cudaStream_t st) {
  if (M <= 0 || N <= 0) return DSVGP_OK;
  if (!Ah || !Al || !Bh || !Bl || !C || K <= 0 || chunk < 1 || (C2 && !D2) || (C2lo && !C2)) return DSVGP_ERR_ARG;
  if (nsplit > 1 && (!split_ws || C2 || Clo)) return DSVGP_ERR_ARG;
  if (!gemm_tc_supported(Ah, lda, Bh, ldb, b_kmajor, N) || !gemm_tc_supported(Al, lda, Bl, ldb, b_kmajor, N))
    return DSVGP_ERR_ARG;
  const int cg = g_tc_cta_group;
  const int bnl = tc::BN / cg;                       // B columns staged per CTA
  CUtensorMap mAh, mAl, mBh, mBl;
  bool ok = tc::map_kmajor(&mAh, Ah, lda, M, K, tc::BM) && tc::map_kmajor(&mAl, Al, lda, M, K, tc::BM);
  if (b_kmajor) ok = ok && tc::map_kmajor(&mBh, Bh, ldb, N, K, bnl) && tc::map_kmajor(&mBl, Bl, ldb, N, K, bnl);
  else ok = ok && tc::map_mnmajor(&mBh, Bh, ldb, K, N, bnl) && tc::map_mnmajor(&mBl, Bl, ldb, K, N, bnl);
  if (!ok) return DSVGP_ERR_ARG;
  static bool attr_set[64] = {};                    // the attribute is PER DEVICE
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    if (cudaFuncSetAttribute(tc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Geo<1>::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(tc::gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Geo<2>::SMEM_BYTES) != cudaSuccess)
      return DSVGP_ERR_LAUNCH;
    attr_set[dev & 63] = true;
  }
  static bool attr_set_p[64] = {};
  const int mtiles = cg == 2 ? ((ceil_div(M, tc::BM) + 1) & ~1) : ceil_div(M, tc::BM);   // whole CTA pairs
  auto launch = [&](tc::Params pp, int nz) -> int {             // returns the number of split-K slices written
    if (cg == 2 && g_tc_persistent) {
      if (!attr_set_p[dev & 63]) {
        if (cudaFuncSetAttribute(tc::gemm_tc2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::P_SMEM_BYTES) != cudaSuccess) return nz;
        attr_set_p[dev & 63] = true;
      }
      int P = tc::resident_pairs(tc::gemm_tc2p_kernel, dev);
      if (g_tc_max_pairs > 0) P = std::min(P, g_tc_max_pairs);
      const int* offs; const int4* items; int nz_eff = nz;
      if (tc::get_sched(tc::SchedKey{dev, M, N, K, a_tri, c_lower, nz, P, tc::BK}, st, &offs, &items, &nz_eff)) {
        pp.nz = nz; pp.trace = g_tc_trace; pp.trace_cap = g_tc_trace_cap;
        tc::gemm_tc2p_kernel<<<dim3(2 * P), tc::P_THREADS, tc::P_SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp, offs, items);
        return nz_eff;
      }
    }
    dim3 grid(mtiles, ceil_div(N, tc::BN), nz);
    if (cg == 2) tc::gemm_tc2_kernel<<<grid, tc::THREADS, tc::Geo<2>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    else tc::gemm_tc_kernel<<<grid, tc::THREADS, tc::Geo<1>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    return nz;
  };
  if (nsplit > 1) {
    const int64_t ldp = round_up64(N, 4), stride = (int64_t)M * ldp;
    tc::Params p{split_ws, nullptr, nullptr, nullptr, nullptr, nullptr, ldp, 0, 0, 0, stride, M, N, K, 1.f, 0.f,
                 a_tri, c_lower, b_kmajor, chunk, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
    const int nslices = launch(p, nsplit);
    CHECK_LAUNCH();
    dim3 rgrid(ceil_div(N, 256), M);
    tc::splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(split_ws, nslices, stride, ldp, M, N, alpha, beta, D, ldd, C, ldc, c_lower,
                                                    nullptr);
    CHECK_LAUNCH();
    return DSVGP_OK;
  }
  tc::Params p{C, D, C2, D2, Clo, C2lo, ldc, ldd, ldc2, ldd2, 0, M, N, K, alpha, beta, a_tri, c_lower, b_kmajor, chunk,
               nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
  launch(p, 1);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int gemm_tch_supported(const void* A, int64_t lda, const void* B, int64_t ldb, int b_kmajor, int N) {
  if (!tc::encode_fn()) return 0;
  if ((lda & 7) || (ldb & 7)) return 0;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return 0;
  if (!b_kmajor && ldb < (int64_t)((N + 63) / 64) * 64) return 0;     // the last 64-column block must stay inside the row
  return 1;
}

int gemm_tch(const void* Ah_, const void* Al_, int64_t lda, const void* Bh_, const void* Bl_, int64_t ldb, int b_kmajor, int M,
             int N, int K, float alpha, float beta, const float* ab_inv, float* C, int64_t ldc, const float* D, int64_t ldd,
             float* C2, int64_t ldc2, const float* D2, int64_t ldd2, void* Ch, void* Cl, int64_t ldch, const float* c_scale,
             void* C2h, void* C2l, int64_t ldc2h, const float* c2_scale, int a_tri, int c_lower, int chunk, int nsplit,
             float* split_ws, cudaStream_t st) {
  if (M <= 0 || N <= 0) return DSVGP_OK;
  const __half *Ah = static_cast<const __half*>(Ah_), *Al = static_cast<const __half*>(Al_);
  const __half *Bh = static_cast<const __half*>(Bh_), *Bl = static_cast<const __half*>(Bl_);
  if (!Ah || !Al || !Bh || !Bl || !ab_inv || K <= 0 || chunk < 1) return DSVGP_ERR_ARG;
  if (!C && !Ch && !C2h) return DSVGP_ERR_ARG;
  if ((C2 || C2h) && !D2) return DSVGP_ERR_ARG;
  if ((Ch && (!Cl || !c_scale)) || (C2h && (!C2l || !c2_scale))) return DSVGP_ERR_ARG;
  if (beta != 0.f && !D && !C) return DSVGP_ERR_ARG;
  if (nsplit > 1 && (!split_ws || !C || C2 || Ch || C2h)) return DSVGP_ERR_ARG;
  if (!gemm_tch_supported(Ah, lda, Bh, ldb, b_kmajor, N) || !gemm_tch_supported(Al, lda, Bl, ldb, b_kmajor, N))
    return DSVGP_ERR_ARG;
  const int cg = g_tc_cta_group;
  const int bnt = (cg == 2 && g_tc_tile_n == 128) ? 128 : tc::BN;
  const int bnl = bnt / cg;
  CUtensorMap mAh, mAl, mBh, mBl;
  bool ok = tc::map_kmajor_h(&mAh, Ah, lda, M, K, tc::BM) && tc::map_kmajor_h(&mAl, Al, lda, M, K, tc::BM);
  if (b_kmajor) ok = ok && tc::map_kmajor_h(&mBh, Bh, ldb, N, K, bnl) && tc::map_kmajor_h(&mBl, Bl, ldb, N, K, bnl);
  else ok = ok && tc::map_mnmajor_h(&mBh, Bh, ldb, K, N, bnl) && tc::map_mnmajor_h(&mBl, Bl, ldb, K, N, bnl);
  if (!ok) return DSVGP_ERR_ARG;
  static bool attr_set[64] = {};                    // the attribute is PER DEVICE
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    if (cudaFuncSetAttribute(tc::gemm_tch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Geo<1>::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(tc::gemm_tch2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Geo<2>::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(tc::gemm_tch2n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Geo<2, 128>::SMEM_BYTES) != cudaSuccess)
      return DSVGP_ERR_LAUNCH;
    attr_set[dev & 63] = true;
  }
  const int mtiles = cg == 2 ? ((ceil_div(M, tc::BM) + 1) & ~1) : ceil_div(M, tc::BM);
  static bool attr_set_p[64] = {};
  auto launch = [&](tc::Params pp, int nz) -> int {             // returns the number of split-K slices written
    if (cg == 2 && bnt == tc::BN && g_tc_persistent) {
      if (!attr_set_p[dev & 63]) {
        if (cudaFuncSetAttribute(tc::gemm_tch2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::P_SMEM_BYTES) != cudaSuccess) return nz;
        attr_set_p[dev & 63] = true;
      }
      int P = tc::resident_pairs(tc::gemm_tch2p_kernel, dev);
      if (g_tc_max_pairs > 0) P = std::min(P, g_tc_max_pairs);
      const int* offs; const int4* items; int nz_eff = nz;
      if (tc::get_sched(tc::SchedKey{dev, M, N, K, a_tri, c_lower, nz, P, tc::BKH}, st, &offs, &items, &nz_eff)) {
        pp.nz = nz; pp.trace = g_tc_trace; pp.trace_cap = g_tc_trace_cap;
        tc::gemm_tch2p_kernel<<<dim3(2 * P), tc::P_THREADS, tc::P_SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp, offs, items);
        return nz_eff;
      }
    }
    dim3 grid(mtiles, ceil_div(N, bnt), nz);
    if (bnt == 128) tc::gemm_tch2n_kernel<<<grid, tc::Geo<2, 128>::NTHREADS, tc::Geo<2, 128>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    else if (cg == 2) tc::gemm_tch2_kernel<<<grid, tc::THREADS, tc::Geo<2>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    else tc::gemm_tch_kernel<<<grid, tc::THREADS, tc::Geo<1>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    return nz;
  };
  if (nsplit > 1) {
    const int64_t ldp = round_up64(N, 4), stride = (int64_t)M * ldp;
    tc::Params p{split_ws, nullptr, nullptr, nullptr, nullptr, nullptr, ldp, 0, 0, 0, stride, M, N, K, 1.f, 0.f,
                 a_tri, c_lower, b_kmajor, chunk, ab_inv, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
    const int nslices = launch(p, nsplit);
    CHECK_LAUNCH();
    dim3 rgrid(ceil_div(N, 256), M);
    tc::splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(split_ws, nslices, stride, ldp, M, N, alpha, beta, D, ldd, C, ldc, c_lower,
                                                    ab_inv);
    CHECK_LAUNCH();
    return DSVGP_OK;
  }
  tc::Params p{C, D, C2, D2, nullptr, nullptr, ldc, ldd, ldc2, ldd2, 0, M, N, K, alpha, beta, a_tri, c_lower, b_kmajor, chunk,
               ab_inv, c_scale, c2_scale, static_cast<__half*>(Ch), static_cast<__half*>(Cl), static_cast<__half*>(C2h),
               static_cast<__half*>(C2l), ldch, ldc2h};
  launch(p, 1);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int split_lo(const float* x, int64_t ldx, float* lo, int64_t ldl, int rows, int cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(ceil_div(cols, 4), 256), rows);
  tc::split_lo_kernel<<<grid, 256, 0, st>>>(x, ldx, lo, ldl, rows, cols);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

int transpose_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return DSVGP_OK;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  tc::transpose_kernel<<<grid, block, 0, st>>>(src, lds, dst, ldd, rows, cols);
  CHECK_LAUNCH();
  return DSVGP_OK;
}

}  // namespace dsvgp
