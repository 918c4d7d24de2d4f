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
#include <cuda.h>
#include <cuda_fp16.h>

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
  const int mtiles = cg == 2 ? ((ceil_div(M, tc::BM) + 1) & ~1) : ceil_div(M, tc::BM);   // whole CTA pairs
  auto launch = [&](const tc::Params& pp, int nz) {
    dim3 grid(mtiles, ceil_div(N, tc::BN), nz);
    if (cg == 2) tc::gemm_tc2_kernel<<<grid, tc::THREADS, tc::Geo<2>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    else tc::gemm_tc_kernel<<<grid, tc::THREADS, tc::Geo<1>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
  };
  if (nsplit > 1) {
    const int64_t ldp = round_up64(N, 4), stride = (int64_t)M * ldp;
    tc::Params p{split_ws, nullptr, nullptr, nullptr, nullptr, nullptr, ldp, 0, 0, 0, stride, M, N, K, 1.f, 0.f,
                 a_tri, c_lower, b_kmajor, chunk, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
    launch(p, nsplit);
    CHECK_LAUNCH();
    dim3 rgrid(ceil_div(N, 256), M);
    tc::splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(split_ws, nsplit, stride, ldp, M, N, alpha, beta, D, ldd, C, ldc, c_lower,
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
  auto launch = [&](const tc::Params& pp, int nz) {
    dim3 grid(mtiles, ceil_div(N, bnt), nz);
    if (bnt == 128) tc::gemm_tch2n_kernel<<<grid, tc::Geo<2, 128>::NTHREADS, tc::Geo<2, 128>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    else if (cg == 2) tc::gemm_tch2_kernel<<<grid, tc::THREADS, tc::Geo<2>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
    else tc::gemm_tch_kernel<<<grid, tc::THREADS, tc::Geo<1>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, pp);
  };
  if (nsplit > 1) {
    const int64_t ldp = round_up64(N, 4), stride = (int64_t)M * ldp;
    tc::Params p{split_ws, nullptr, nullptr, nullptr, nullptr, nullptr, ldp, 0, 0, 0, stride, M, N, K, 1.f, 0.f,
                 a_tri, c_lower, b_kmajor, chunk, ab_inv, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
    launch(p, nsplit);
    CHECK_LAUNCH();
    dim3 rgrid(ceil_div(N, 256), M);
    tc::splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(split_ws, nsplit, stride, ldp, M, N, alpha, beta, D, ldd, C, ldc, c_lower,
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
