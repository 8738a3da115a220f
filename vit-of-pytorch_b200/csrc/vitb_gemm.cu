// vitb_gemm.cu — the dense contraction of the ViT encoder path on sm_100a.
//
//   D[M,N] = epilogue( sum_seg A_seg[M,K_seg] * B_seg[N,K_seg]^T )        (contract: vitb200.h)
//
// Design (B200-first, not a port of anything in the reference — the reference only calls ATen):
//   * persistent CTAs (one per SM), static tile scheduler, 384 threads, warp-specialised:
//       warp 0    : TMA producer (cp.async.bulk.tensor, SWIZZLE_128B) into a 4/6-stage smem ring
//       warp 1    : one elected thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM)
//       warp 2    : TMEM allocator (2 accumulator stages so the epilogue overlaps the next tile)
//       warps 4-11: epilogue (two per TMEM lane quadrant, each draining half of the tile's columns in 32 x 32
//                   chunks) — tcgen05.ld TMEM->registers, then one of two families:
//                   (a) register layout (thread = accumulator row): bias / GELU + GELU' on packed fp32 pairs /
//                       x aux, bf16 tiles through 64B-swizzled staging and TMA stores;
//                   (b) staged: transpose through padded fp32 smem so that the side operand (residual, aux) and
//                       the fp32 / accumulating outputs move in row-coalesced 16-byte pieces
//   * K-major and MN-major operands are both fed straight from their HBM layout through the UMMA
//     shared-memory descriptors (no transposition pass for dgrad / wgrad / LinearGeneral weights)
//   * up to 3 (A,B,K) segments accumulate in the same TMEM tile: LoRA rank-r update, bf16x3 split
//   * split-K with fp32 red.global accumulation for the weight-gradient shapes (few output tiles)
//
// Algorithmic work per launch: 2*M*N*sum(K_seg) FLOP; bytes >= 2*(M*K + N*K) + out.
#include <stdlib.h>

#include "../../include/vitb200.h"
#include "vitb_common.cuh"

// Ablation build (vit-of-pytorch_b200/build.py --tools compiles this file a second time with -DVITB_GEMM_DIAG=1 into the
// entry points vitb_gemm_diag / vitb_gemm_diag_mask of a SEPARATE diagnostics library, libvitb200_tools.so): a bit mask
// switches parts of the epilogue OFF so that one GPU call can time the kernel without them (tools/epi_ablate.py).
// Results are then WRONG by construction; the product library contains none of this.
//   1 no TMA store issue   2 no epilogue math   4 no TMEM load   8 no staging-tile writes   16 no side / bias loads
//   32 no per-chunk work at all (handshakes only: the bare mainloop)   64 no async-proxy fence   128 no column sums
#ifdef VITB_GEMM_DIAG
__device__ __constant__ int g_diag_mask;
#define VITB_DIAG(bit) ((g_diag_mask & (bit)) != 0)
#define VITB_GEMM_ENTRY vitb_gemm_diag
#else
#define VITB_DIAG(bit) false
#define VITB_GEMM_ENTRY vitb_gemm
#endif

namespace {

using namespace vitb;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;                         // 2 per TMEM lane quadrant, each owning half of the columns
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int A_BYTES = BM * BK * 2;                 // 16 KiB
constexpr int kStgStride = 34;                       // floats per staged row: 8B-aligned, conflict-free float2 access
constexpr int kStagingFloats = 32 * kStgStride;      // per epilogue warp
constexpr int kStagingBytes = kEpiWarps * kStagingFloats * 4;

// Two experiments of round 2 changed this configuration and were removed again because they bought nothing (A/B files
// under profiles/): 16 epilogue warps instead of 8 (gemm_ew16_r02.txt) and four TMA-store tiles per warp paid for with the
// fourth mainloop stage (gemm_deepstore_r02.txt).  The bf16 epilogues behind a K = 768 mainloop are therefore limited
// neither by warp-level latency hiding nor by the depth of their store pipeline; what they share with the mainloop is the
// L2: a 128 x 256 tile pulls 590 KB of operands for 50 MFLOP, and every output byte competes with that stream.
template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  // no alignment slack: the dynamic shared window starts 1024-byte aligned (checked at kernel entry)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kStagingBytes + 256 /*barriers*/;
};

struct GemmDev {
  int M, N;
  int nseg;
  int kblocks[3];
  int total_kblocks;
  int split_k;
  int m_tiles, n_tiles;
  int epilogue;
  void* D;
  long long ldd;
  int d_bf16;
  int accumulate;
  void* D2;
  long long ldd2;
  const float* bias;
  const float* row_bias;
  int row_bias_group;
  int row_remap_group;
  const void* residual;
  long long ldr;
  int r_bf16;
  const void* aux;
  long long ldaux;
  int vec_ok;  // 128-bit epilogue path usable (set by the host from shapes and alignment)
  float* colsum;  // optional [N]: += column sums of the values stored to D (vectorised path only)
  int tma_store;     // bf16 outputs leave through cp.async.bulk.tensor stores (tmD / tmD2 are valid)
  int d2_grad;       // VITB_EPI_GELU_DG: D2 receives gelu'(v) instead of v
  int aux_grad;      // VITB_EPI_MUL_AUX: aux already holds gelu'(z); the epilogue only multiplies
  int packed_epi;    // GELU + GELU' epilogue on packed fp32 pairs with the bias staged in shared memory (VITB_EPI_PACKED)
  int rowmul;        // VITB_EPI_MUL_AUX in the TMEM register layout: aux rows prefetched a chunk ahead, TMA stores (VITB_EPI_ROWMUL)
  const int* m_dev;  // optional device scalar: rows that hold data (tiles past it are skipped); nullptr = M
  int group_cols;    // Ng of a column-grouped B (merged q|k|v): tile column n0 -> group n0 / Ng; 0 = one group
  long long d_gs;    // distance (elements) between the groups of a grouped fp32 accumulate output; 0 = contiguous D
};

struct TileCoord {
  int m_blk, n_blk, g0, g1;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmDev& p, int tile) {
  TileCoord t;
  const int split = tile % p.split_k;
  const int t2 = tile / p.split_k;
  t.n_blk = t2 % p.n_tiles;
  t.m_blk = t2 / p.n_tiles;
  t.g0 = static_cast<int>((static_cast<long long>(split) * p.total_kblocks) / p.split_k);
  t.g1 = static_cast<int>((static_cast<long long>(split + 1) * p.total_kblocks) / p.split_k);
  return t;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {  // two 8-byte loads (rows are 8B-aligned)
  float4 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.z), "=f"(v.w) : "r"(addr + 8) : "memory");
  return v;
}
// erf-GELU and its derivative from ONE rcp + ONE ex2 (Abramowitz-Stegun 7.1.26, |erf error| < 1.5e-7):
// used where the result is rounded to bf16 anyway; fp32 outputs keep erff.
__device__ __forceinline__ void gelu_fast(float z, float& g, float& dg) {
  const float u = fabsf(z) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, u, 1.0f));
  const float e = ex2_approx(-0.72134752044448170368f * z * z);     // exp(-z^2/2) = exp(-u^2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);                   // erf(|u|)
  const float cdf = 0.5f * (1.0f + copysignf(erf_abs, z));
  g = z * cdf;
  dg = fmaf(z * 0.39894228040143267794f, e, cdf);
}
// Forward only, 13 instructions: the 0.5 of the cdf is folded into the polynomial and the sign into |z|:
//   gelu(z) = 0.5 z + |z| * (0.5 erf(|z|/sqrt2)),  0.5 erf(u) = 0.5 - (0.5 poly(t) t) exp(-u^2)
__device__ __forceinline__ float gelu_fast_fwd(float z) {
  const float az = fabsf(z);
  const float t = rcp_approx(fmaf(0.23164189f, az, 1.0f));          // 1 / (1 + 0.3275911 |z| / sqrt2)
  const float w = 0.84932180f * z;                                   // sqrt(log2(e) / 2) z
  const float e = ex2_approx(-w * w);                                // exp(-z^2 / 2)
  float poly = fmaf(0.5307027145f, t, -0.7265760135f);
  poly = fmaf(poly, t, 0.7107068705f);
  poly = fmaf(poly, t, -0.142248368f);
  poly = fmaf(poly, t, 0.127414796f);
  const float half_erf = fmaf(-poly * t, e, 0.5f);                   // 0.5 erf(|z| / sqrt2)
  return fmaf(az, half_erf, 0.5f * z);
}
// Forward value and derivative together (17 instructions): used when the forward stores gelu'(z) for the
// backward (VITB_EPI_GELU_DG) instead of the pre-activation z.
__device__ __forceinline__ void gelu_fast_both(float z, float& g, float& dg) {
  const float az = fabsf(z);
  const float t = rcp_approx(fmaf(0.23164189f, az, 1.0f));
  const float w = 0.84932180f * z;
  const float e = ex2_approx(-w * w);                                // exp(-z^2 / 2)
  float poly = fmaf(0.5307027145f, t, -0.7265760135f);
  poly = fmaf(poly, t, 0.7107068705f);
  poly = fmaf(poly, t, -0.142248368f);
  poly = fmaf(poly, t, 0.127414796f);
  const float half_erf = fmaf(-poly * t, e, 0.5f);                   // 0.5 erf(|z| / sqrt2)
  g = fmaf(az, half_erf, 0.5f * z);
  dg = fmaf(z * 0.39894228040143267794f, e, 0.5f + copysignf(half_erf, z));   // Phi(z) + z phi(z)
}
// The same value / derivative pair for TWO pre-activations at once on packed fp32 pairs (FFMA2 / FMUL2 / FADD2):
// 13 FMA-pipe issue slots per pair instead of 30, 4 LOP3 on the ALU pipe, 4 MUFU.  gelu(z) = z Phi(z) here
// (Phi is needed for the derivative anyway), the polynomial carries its sign in the coefficients.
__device__ __forceinline__ void gelu_fast_both2(float z0, float z1, uint64_t& g, uint64_t& dg) {
  const uint64_t z = pk2(z0, z1);
  const uint64_t az = pk2(fabsf(z0), fabsf(z1));
  float ti0, ti1;
  upk2(fma2(az, pk2(0.23164189f), pk2(1.0f)), ti0, ti1);
  const uint64_t t = pk2(rcp_approx(ti0), rcp_approx(ti1));       // 1 / (1 + 0.3275911 |z| / sqrt2)
  float ei0, ei1;
  upk2(mul2(mul2(z, z), pk2(-0.72134752044448170368f)), ei0, ei1);
  const uint64_t e = pk2(ex2_approx(ei0), ex2_approx(ei1));       // exp(-z^2 / 2)
  uint64_t np = fma2(pk2(-0.5307027145f), t, pk2(0.7265760135f)); // -(0.5 poly(t))
  np = fma2(np, t, pk2(-0.7107068705f));
  np = fma2(np, t, pk2(0.142248368f));
  np = fma2(np, t, pk2(-0.127414796f));
  float he0, he1;
  upk2(fma2(mul2(np, t), e, pk2(0.5f)), he0, he1);                // 0.5 erf(|z| / sqrt2) >= 0 up to 1.5e-7
  const uint64_t phi = add2(pk2(copysignf(he0, z0), copysignf(he1, z1)), pk2(0.5f));   // Phi(z)
  g = mul2(z, phi);
  dg = fma2(mul2(z, pk2(0.39894228040143267794f)), e, phi);      // Phi(z) + z phi(z)
}
__device__ __forceinline__ uint32_t pack_bf16x2_pair(uint64_t v) {
  float lo, hi;
  upk2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float4 ld_shared_f4_16(uint32_t addr) {   // one 16-byte load (16-byte aligned address)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_shared_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// Vectorised epilogue of one staged 32x32 fp32 sub-tile: lane -> (row%4 = lane/8, 4 columns = lane%8),
// 8 steps of 4 rows; every global access is a 16-byte (fp32) or 8-byte (bf16) piece of a row segment
// that the 8 lanes of a row cover contiguously.  All side loads of the 8 steps are issued up front.
//   RES: 0 none, 1 fp32 residual, 2 bf16 residual.   ACC: fp32 red.global accumulation.
// FULL: all 32 rows and all 32 columns of the chunk are inside the matrix (the common case: no per-row
// predicates, running row pointers instead of a 64-bit multiply per access).
// Side operand of the staged epilogue (the residual, or the aux tensor of GELU_BWD / MUL_AUX), kept in its raw
// form: lane (row%4, 4 columns) holds 8 bytes (bf16) or 16 bytes (fp32) for each of its 8 rows.
#pragma nv_diag_suppress 177   // members unused in the instantiations without a side operand
template <bool OUT_BF16, int EPI, int RES>
struct SideOf {
  static constexpr bool HAS = (RES != 0 || EPI == VITB_EPI_GELU_BWD);
  static constexpr bool BF16 = (EPI == VITB_EPI_GELU_BWD) ? OUT_BF16 : (RES == 2);
  static constexpr int ESIZE = BF16 ? 2 : 4;
  __device__ static __forceinline__ const void* base(const GemmDev& p) { return (EPI == VITB_EPI_GELU_BWD) ? p.aux : p.residual; }
  __device__ static __forceinline__ long long ld(const GemmDev& p) { return (EPI == VITB_EPI_GELU_BWD) ? p.ldaux : p.ldr; }
  __device__ static __forceinline__ uint4 load(const char* sp) {
    if constexpr (BF16) {
      const uint2 u = *reinterpret_cast<const uint2*>(sp);
      return make_uint4(u.x, u.y, 0u, 0u);
    } else {
      return *reinterpret_cast<const uint4*>(sp);
    }
  }
  __device__ static __forceinline__ float4 value(const uint4& r) {
    if constexpr (BF16) return make_float4(bf16_lo(r.x), bf16_hi(r.x), bf16_lo(r.y), bf16_hi(r.y));
    else return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
  }
};
#pragma nv_diag_default 177

// All 8 side loads of one chunk (predicated; absent elements read as zero).  Used for the first chunk of a tile,
// BEFORE the accumulator barrier is waited on: the side operand does not depend on the MMA.
template <bool OUT_BF16, int EPI, int RES>
__device__ __forceinline__ void side_fetch(const GemmDev& p, int lane, int row_base, int col0, bool lead_split,
                                           uint4 (&raw)[8]) {
  using S = SideOf<OUT_BF16, EPI, RES>;
  if constexpr (S::HAS) {
    const int rsub = lane >> 3;
    const int col = col0 + (lane & 7) * 4;
    const bool want = ((EPI == VITB_EPI_GELU_BWD) || lead_split) && col < p.N;
    const long long lds = S::ld(p);
    const char* sp = reinterpret_cast<const char*>(S::base(p)) + (static_cast<long long>(row_base + rsub) * lds + col) * S::ESIZE;
    const long long sstep = 4 * lds * S::ESIZE;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      raw[it] = make_uint4(0u, 0u, 0u, 0u);
      if (want && row_base + it * 4 + rsub < p.M) raw[it] = S::load(sp);
      sp += sstep;
    }
  }
}

// `raw` holds this chunk's side operand on entry.  While the chunk is processed, each consumed slot is refilled
// with the same rows of the NEXT chunk (columns next_col0 .., or none when next_col0 < 0): the loads fly during
// the rest of this chunk and the next chunk's TMEM read / staging, with no extra registers — the eight warps of
// the epilogue cannot hide HBM latency by occupancy (round-1 ncu: GELU' GEMM at 31 % issue-active, 2.3 TB/s).
template <bool OUT_BF16, int EPI, int RES, bool ACC, bool FULL>
__device__ __forceinline__ void epi_vec_body(const GemmDev& p, uint32_t stg, int lane, int row_base, int col0,
                                             bool lead_split, uint4 (&raw)[8], int next_col0, long long d_off) {
  using S = SideOf<OUT_BF16, EPI, RES>;
  const int rsub = lane >> 3;
  const int c4 = (lane & 7) * 4;
  const int col = col0 + c4;
  const bool col_ok = FULL || col < p.N;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias != nullptr && lead_split && col_ok) bv = *reinterpret_cast<const float4*>(p.bias + col);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);   // column sums of the stored values (bias gradients)
  const long long r0 = static_cast<long long>(row_base + rsub);
  const char* spn = nullptr;      // this lane's first row of the next chunk's side operand
  long long sstep = 0;
  bool next_ok = false;
  if constexpr (S::HAS) {
    next_ok = next_col0 >= 0 && ((EPI == VITB_EPI_GELU_BWD) || lead_split) && (next_col0 + c4 < p.N);
    const long long lds = S::ld(p);
    spn = reinterpret_cast<const char*>(S::base(p)) + (r0 * lds + (next_col0 + c4)) * S::ESIZE;
    sstep = 4 * lds * S::ESIZE;
  }
  char* dp = reinterpret_cast<char*>(p.D) + (r0 * p.ldd + col + d_off) * (OUT_BF16 ? 2 : 4);   // d_off: grouped output
  const long long dstep = 4 * p.ldd * (OUT_BF16 ? 2 : 4);
  char* d2p = nullptr;
  long long d2step = 0;
  if constexpr (EPI == VITB_EPI_GELU) {
    if (p.D2 != nullptr) {
      d2p = reinterpret_cast<char*>(p.D2) + (r0 * p.ldd2 + col) * (OUT_BF16 ? 2 : 4);
      d2step = 4 * p.ldd2 * (OUT_BF16 ? 2 : 4);
    }
  }
  uint32_t sa = stg + static_cast<uint32_t>(rsub * (kStgStride * 4) + c4 * 4);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (col_ok && (FULL || row_base + it * 4 + rsub < p.M)) {
      float4 v = ld_shared_f4(sa);
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      if constexpr (EPI == VITB_EPI_GELU) {
        float4 s2 = v;   // what D2 receives: the pre-activation, or gelu'(pre-activation)
        if constexpr (OUT_BF16) {
          if (p.d2_grad) {
            gelu_fast_both(v.x, v.x, s2.x); gelu_fast_both(v.y, v.y, s2.y);
            gelu_fast_both(v.z, v.z, s2.z); gelu_fast_both(v.w, v.w, s2.w);
          } else {
            v.x = gelu_fast_fwd(v.x); v.y = gelu_fast_fwd(v.y); v.z = gelu_fast_fwd(v.z); v.w = gelu_fast_fwd(v.w);
          }
        } else {
          if (p.d2_grad) s2 = make_float4(gelu_erf_grad(v.x), gelu_erf_grad(v.y), gelu_erf_grad(v.z), gelu_erf_grad(v.w));
          v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        }
        if (d2p != nullptr) {
          if constexpr (OUT_BF16) {
            uint2 z;
            z.x = pack_bf16x2(s2.x, s2.y); z.y = pack_bf16x2(s2.z, s2.w);
            *reinterpret_cast<uint2*>(d2p) = z;
          } else {
            *reinterpret_cast<float4*>(d2p) = s2;
          }
        }
      } else if constexpr (EPI == VITB_EPI_GELU_BWD) {
        const float4 sd = S::value(raw[it]);
        if (p.aux_grad) {
          v.x *= sd.x; v.y *= sd.y; v.z *= sd.z; v.w *= sd.w;
        } else if constexpr (OUT_BF16) {
          float g, d0, d1, d2, d3;
          gelu_fast(sd.x, g, d0); gelu_fast(sd.y, g, d1);
          gelu_fast(sd.z, g, d2); gelu_fast(sd.w, g, d3);
          v.x *= d0; v.y *= d1; v.z *= d2; v.w *= d3;
        } else {
          v.x *= gelu_erf_grad(sd.x); v.y *= gelu_erf_grad(sd.y);
          v.z *= gelu_erf_grad(sd.z); v.w *= gelu_erf_grad(sd.w);
        }
      }
      if constexpr (RES != 0) {
        const float4 sd = S::value(raw[it]);
        v.x += sd.x; v.y += sd.y; v.z += sd.z; v.w += sd.w;
      }
      cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
      if constexpr (OUT_BF16) {
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dp) = o;
      } else {
        if constexpr (ACC) atomicAdd(reinterpret_cast<float4*>(dp), v);
        else *reinterpret_cast<float4*>(dp) = v;
      }
    }
    if constexpr (S::HAS) {   // slot `it` is consumed: refill it with the same row of the next chunk
      if (next_ok && (FULL || row_base + it * 4 + rsub < p.M)) raw[it] = S::load(spn);
      spn += sstep;
    }
    sa += 4 * (kStgStride * 4);
    dp += dstep;
    if constexpr (EPI == VITB_EPI_GELU) d2p += d2step;
  }
  if (p.colsum != nullptr) {  // warp-uniform: fold the 4 row groups (lanes l, l+8, l+16, l+24), one red per column
    cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8);  cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
    cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8);  cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
    cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
    cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
    if (lane < 8 && col_ok) atomicAdd(reinterpret_cast<float4*>(p.colsum + col), cs);
  }
}

template <bool OUT_BF16, int EPI, int RES, bool ACC>
__device__ __forceinline__ void epi_vec(const GemmDev& p, uint32_t stg, int lane, int row_base, int col0,
                                        bool lead_split, uint4 (&raw)[8], int next_col0, long long d_off = 0) {
  if (row_base + 32 <= p.M && col0 + 32 <= p.N)   // warp-uniform
    epi_vec_body<OUT_BF16, EPI, RES, ACC, true>(p, stg, lane, row_base, col0, lead_split, raw, next_col0, d_off);
  else
    epi_vec_body<OUT_BF16, EPI, RES, ACC, false>(p, stg, lane, row_base, col0, lead_split, raw, next_col0, d_off);
}


// ---- bf16 outputs through TMA stores -------------------------------------------------------------------
// bf16 outputs without residual / column sums do their element-wise math in the TMEM register layout (thread =
// row, 32 columns).  The first version of this path staged packed bf16 through padded shared memory and wrote
// 8-byte row pieces with ld.shared + st.global; its SASS spent ~15 instructions per store on 64-bit row
// addressing and bounds predicates (profiles/ncu_gemm_r01.txt).  Here the thread that owns a row writes its 32
// packed values (64 B) into a 64B-swizzled 32 x 32 staging tile with four conflict-free 16-byte shared stores,
// and one lane hands the tile to the TMA unit, which does the addressing, the coalescing and the M / N
// clipping.  Two tiles per warp alternate, so a store is only waited for when its tile comes up for reuse two
// stores later.  Measured (profiles/gemm_bench_r01b.txt): fc1+GELU 0.167 -> 0.132 ms, fc1+bias 0.114 -> 0.105 ms.
constexpr int kTmaTileBytes = 32 * 64;

// acquire the next staging tile of this warp: the store that last read it (NT stores ago) has drained
template <int NT>
__device__ __forceinline__ uint32_t tma_tile_acquire(uint32_t tbuf, int& which, int lane) {
  const uint32_t buf = tbuf + static_cast<uint32_t>(which) * kTmaTileBytes;
  which = (which + 1 == NT) ? 0 : which + 1;
  if (lane == 0) bulk_wait_read<NT - 1>();
  __syncwarp();
  return buf;
}
// this thread's row: 16-byte unit j (columns 8j .. 8j+7) of the 64B-swizzled tile
__device__ __forceinline__ void tma_tile_write_unit(uint32_t buf, int lane, int j, float a0, float a1, float a2, float a3,
                                                    float a4, float a5, float a6, float a7) {
  const uint32_t x = (static_cast<uint32_t>(lane) >> 1) & 3u;   // SWIZZLE_64B: 16-byte unit ^= address bits [7,9)
  if (VITB_DIAG(8)) return;
  st_shared_v4(buf + static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(j) ^ x) << 4), pack_bf16x2(a0, a1),
               pack_bf16x2(a2, a3), pack_bf16x2(a4, a5), pack_bf16x2(a6, a7));
}
__device__ __forceinline__ void tma_tile_release(const CUtensorMap* tm, uint32_t buf, int lane, int row_base, int col0) {
  if (!VITB_DIAG(64)) fence_proxy_async_smem();   // generic-proxy writes -> visible to the async proxy (TMA)
  __syncwarp();
  if (lane == 0 && !VITB_DIAG(1)) {
    tma_store_2d(tm, buf, col0, row_base);
    bulk_commit();
  }
}
template <int NT>
__device__ __forceinline__ void tma_store_rows_bf16(const CUtensorMap* tm, uint32_t tbuf, int& which, int lane,
                                                    const float (&v)[32], int row_base, int col0) {
  const uint32_t buf = tma_tile_acquire<NT>(tbuf, which, lane);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    tma_tile_write_unit(buf, lane, j, v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3], v[8 * j + 4], v[8 * j + 5],
                        v[8 * j + 6], v[8 * j + 7]);
  tma_tile_release(tm, buf, lane, row_base, col0);
}

// One 32-row x 32-column chunk of a bf16 output: bias / GELU in the TMEM register layout, then TMA stores.
template <int EPI, int NT>
__device__ __forceinline__ void epi_rows_bf16_tma(const GemmDev& p, const CUtensorMap* tmD, const CUtensorMap* tmD2,
                                                  uint32_t tbuf, int& which, int lane, int row_base, int col0,
                                                  const uint32_t (&r)[32]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      if (col0 + 4 * j4 < p.N) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j4);
        v[4 * j4] += b.x; v[4 * j4 + 1] += b.y; v[4 * j4 + 2] += b.z; v[4 * j4 + 3] += b.w;
      }
    }
  }
  if constexpr (EPI == VITB_EPI_GELU) {
    if (p.D2 != nullptr && p.d2_grad) {
      // D2 = gelu'(v), D = gelu(v): the derivative goes to its tile 8 columns at a time while v becomes gelu(v)
      const uint32_t buf = tma_tile_acquire<NT>(tbuf, which, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) gelu_fast_both(v[8 * j + i], v[8 * j + i], d[i]);
        tma_tile_write_unit(buf, lane, j, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
      }
      tma_tile_release(tmD2, buf, lane, row_base, col0);
    } else {
      if (p.D2 != nullptr) tma_store_rows_bf16<NT>(tmD2, tbuf, which, lane, v, row_base, col0);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_fast_fwd(v[j]);
    }
  }
  tma_store_rows_bf16<NT>(tmD, tbuf, which, lane, v, row_base, col0);
}

// GELU + GELU' chunk on packed pairs.  The 32 bias values of the chunk sit in this warp's shared-memory slot
// (published by the epilogue loop one chunk ahead: with 227 KB of shared memory in use there is no L1 left, and
// the per-chunk __ldg of the bias paid an L2 round trip in front of the math — profiles/ncu_r01c.txt).
template <int NT>
__device__ __forceinline__ void epi_rows_gelu_dg_packed(const CUtensorMap* tmD, const CUtensorMap* tmD2, uint32_t tbuf,
                                                        int& which, int lane, int row_base, int col0,
                                                        const uint32_t (&r)[32], uint32_t bias_slot, bool has_bias) {
  const uint32_t bufd = tma_tile_acquire<NT>(tbuf, which, lane);   // derivative tile
  uint32_t gq[16];                                               // gelu(v) as packed bf16 pairs, stored after the loop
#pragma unroll
  for (int j = 0; j < 4; ++j) {                                  // 8 columns per 16-byte unit
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]);
    if (has_bias && !VITB_DIAG(16)) {
      const float4 b0 = ld_shared_f4_16(bias_slot + j * 32), b1 = ld_shared_f4_16(bias_slot + j * 32 + 16);
      uint64_t s;
      s = add2(pk2(v[0], v[1]), pk2(b0.x, b0.y)); upk2(s, v[0], v[1]);
      s = add2(pk2(v[2], v[3]), pk2(b0.z, b0.w)); upk2(s, v[2], v[3]);
      s = add2(pk2(v[4], v[5]), pk2(b1.x, b1.y)); upk2(s, v[4], v[5]);
      s = add2(pk2(v[6], v[7]), pk2(b1.z, b1.w)); upk2(s, v[6], v[7]);
    }
    uint32_t dq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint64_t g, dg;
      if (VITB_DIAG(2)) { g = pk2(v[2 * i], v[2 * i + 1]); dg = g; }
      else gelu_fast_both2(v[2 * i], v[2 * i + 1], g, dg);
      gq[4 * j + i] = pack_bf16x2_pair(g);
      dq[i] = pack_bf16x2_pair(dg);
    }
    const uint32_t x = (static_cast<uint32_t>(lane) >> 1) & 3u;   // SWIZZLE_64B, as tma_tile_write_unit
    if (!VITB_DIAG(8))
      st_shared_v4(bufd + static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(j) ^ x) << 4), dq[0], dq[1], dq[2], dq[3]);
  }
  tma_tile_release(tmD2, bufd, lane, row_base, col0);
  const uint32_t bufg = tma_tile_acquire<NT>(tbuf, which, lane);   // value tile
  const uint32_t x = (static_cast<uint32_t>(lane) >> 1) & 3u;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (!VITB_DIAG(8))
      st_shared_v4(bufg + static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(j) ^ x) << 4), gq[4 * j], gq[4 * j + 1],
                   gq[4 * j + 2], gq[4 * j + 3]);
  tma_tile_release(tmD, bufg, lane, row_base, col0);
}

// ---- v *= aux in the TMEM register layout (the fc2 dgrad of the fused block: dz = (dy W2) * gelu'(z)) ------------
// The staged version of this epilogue transposes every 32 x 32 chunk through padded fp32 shared memory so that the
// aux loads and the stores are row-coalesced; its eight warps then sit on a chain of dependent shared-memory round
// trips per chunk (profiles/ncu_r01c.txt: 31 % issue-active, 0.161 ms against 0.095 ms for the bare GEMM).  Here the
// thread that owns accumulator row r reads its own 64 bytes of aux (four 16-byte loads, requested one chunk ahead),
// multiplies in registers, and the bf16 tile leaves through the TMA store path of the plain epilogue.  The column
// sums (bias gradient of fc1) come from a transpose-reduce across the warp: 31 shuffles leave column j's sum over
// the 32 rows in lane j, one 128-byte red per chunk.
__device__ __forceinline__ void aux_rows_load(const GemmDev& p, int lane, int row_base, int col0, uint4 (&dst)[4]) {
  const int row = row_base + lane;
  const char* ap = reinterpret_cast<const char*>(p.aux) + (static_cast<long long>(row) * p.ldaux + col0) * 2;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    dst[u] = make_uint4(0u, 0u, 0u, 0u);
    if (row < p.M && col0 + 8 * u < p.N && !VITB_DIAG(16)) dst[u] = __ldg(reinterpret_cast<const uint4*>(ap) + u);
  }
}
template <int NT>
__device__ __forceinline__ void epi_rows_mul_aux(const GemmDev& p, const CUtensorMap* tmD, uint32_t tbuf, int& which,
                                                 int lane, int row_base, int col0, const uint32_t (&r)[32],
                                                 const uint4 (&ax)[4]) {
  float v[32];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    v[8 * u + 0] = __uint_as_float(r[8 * u + 0]) * bf16_lo(ax[u].x);
    v[8 * u + 1] = __uint_as_float(r[8 * u + 1]) * bf16_hi(ax[u].x);
    v[8 * u + 2] = __uint_as_float(r[8 * u + 2]) * bf16_lo(ax[u].y);
    v[8 * u + 3] = __uint_as_float(r[8 * u + 3]) * bf16_hi(ax[u].y);
    v[8 * u + 4] = __uint_as_float(r[8 * u + 4]) * bf16_lo(ax[u].z);
    v[8 * u + 5] = __uint_as_float(r[8 * u + 5]) * bf16_hi(ax[u].z);
    v[8 * u + 6] = __uint_as_float(r[8 * u + 6]) * bf16_lo(ax[u].w);
    v[8 * u + 7] = __uint_as_float(r[8 * u + 7]) * bf16_hi(ax[u].w);
  }
  if (p.colsum != nullptr && !VITB_DIAG(128)) {   // rows >= M and columns >= N hold exact zeros (zero-filled operands, zeroed aux)
    const float s = warp_transpose_sum32(v, lane);
    if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, s);
  }
  tma_store_rows_bf16<NT>(tmD, tbuf, which, lane, v, row_base, col0);
}

// Scalar epilogue with every option (row bias, patch-embedding row remap, odd widths): lane == column.
__device__ __noinline__ void epi_generic(const GemmDev& p, uint32_t stg, int lane, int row_base, int col0,
                                         bool lead_split) {
  const int col = col0 + lane;
  if (col >= p.N) return;
  const float bv = (p.bias != nullptr && lead_split) ? p.bias[col] : 0.f;
  for (int rr = 0; rr < 32; ++rr) {
    const int m = row_base + rr;
    if (m >= p.M) break;
    float v = ld_shared_f1(stg + rr * (kStgStride * 4) + lane * 4) + bv;
    if (p.row_bias != nullptr && lead_split)
      v += p.row_bias[static_cast<long long>(m / p.row_bias_group) * p.N + col];
    if (p.epilogue == VITB_EPI_GELU) {
      if (p.D2 != nullptr) {
        const float s2 = p.d2_grad ? gelu_erf_grad(v) : v;
        if (p.d_bf16) reinterpret_cast<__nv_bfloat16*>(p.D2)[static_cast<long long>(m) * p.ldd2 + col] = __float2bfloat16(s2);
        else reinterpret_cast<float*>(p.D2)[static_cast<long long>(m) * p.ldd2 + col] = s2;
      }
      v = gelu_erf(v);
    } else if (p.epilogue == VITB_EPI_GELU_BWD) {
      const long long ai = static_cast<long long>(m) * p.ldaux + col;
      const float ax = p.d_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.aux)[ai])
                                : reinterpret_cast<const float*>(p.aux)[ai];
      v *= p.aux_grad ? ax : gelu_erf_grad(ax);
    }
    int om = m, rm = m;
    if (p.row_remap_group > 0) {
      om = m + m / p.row_remap_group + 1;
      rm = m % p.row_remap_group + 1;
    }
    if (p.residual != nullptr && lead_split) {
      const long long ri = static_cast<long long>(rm) * p.ldr + col;
      v += p.r_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[ri])
                    : reinterpret_cast<const float*>(p.residual)[ri];
    }
    if (p.d_bf16) {
      reinterpret_cast<__nv_bfloat16*>(p.D)[static_cast<long long>(om) * p.ldd + col] = __float2bfloat16(v);
    } else {
      float* dst = reinterpret_cast<float*>(p.D) + static_cast<long long>(om) * p.ldd + col;
      if (p.accumulate) atomicAdd(dst, v);
      else *dst = v;
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
vitb_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2,
                 const __grid_constant__ GemmDev p) {
  using C = Cfg<BN>;
  constexpr int NT = 2;                           // TMA-store staging tiles per epilogue warp
  constexpr int CHUNKS = BN / 64;                 // 32-column chunks of a tile drained by one epilogue warp
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) {  // SWIZZLE_128B tiles need 1024-byte aligned stage bases
    if (threadIdx.x == 0) printf("vitb_gemm: dynamic shared memory base 0x%x is not 1024-byte aligned\n", smem_u32(smem));
    __trap();
  }
  float* staging = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + kStagingBytes);
  // bars: [0,STAGES) full, [STAGES,2*STAGES) empty, then tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = smem_u32(bars + C::STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * C::STAGES);
  const uint32_t tempty0 = smem_u32(bars + 2 * C::STAGES + 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (p.nseg > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
    if (p.nseg > 2) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // nothing above touches global memory; everything below may

  // row count from device memory (Res-ViT token compaction): only the tile loop shrinks, every bound stays M
  int m_tiles = p.m_tiles;
  if (p.m_dev != nullptr) {
    const int m_eff = min(max(*p.m_dev, 0), p.M);
    m_tiles = (m_eff + BM - 1) / BM;
  }
  const int total_tiles = m_tiles * p.n_tiles * p.split_k;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int m0 = t.m_blk * BM;
        const int n0 = t.n_blk * BN;
        const int bg = p.group_cols > 0 ? n0 / p.group_cols : 0;     // column group of this tile (tiles never straddle groups)
        const int nloc = n0 - bg * p.group_cols;
        for (int g = t.g0; g < t.g1; ++g) {
          int seg = 0, kb = g;
          if (kb >= p.kblocks[0]) { kb -= p.kblocks[0]; seg = 1; }
          if (seg == 1 && kb >= p.kblocks[1]) { kb -= p.kblocks[1]; seg = 2; }
          const CUtensorMap* ma = seg == 0 ? &tmA0 : (seg == 1 ? &tmA1 : &tmA2);
          const CUtensorMap* mb = seg == 0 ? &tmB0 : (seg == 1 ? &tmB1 : &tmB2);
          mbar_wait(empty0 + 8 * stage, phase ^ 1u);
          const uint32_t fb = full0 + 8 * stage;
          mbar_arrive_expect_tx(fb, C::STAGE_BYTES);
          const uint32_t sA = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sB = sA + A_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(ma, fb, sA, kb * BK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(ma, fb, sA + i * 8192, m0 + 64 * i, kb * BK);
          }
          // B maps are 3-D: (k, n within the column group, group); a plain GEMM has one group
          if constexpr (!B_MN) {
            tma_load_3d(mb, fb, sB, kb * BK, nloc, bg);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_3d(mb, fb, sB + i * 8192, nloc + 64 * i, kb * BK, bg);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int g = t.g0; g < t.g1; ++g) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc_sw128(sA + k * 2048, BK * 128, 1024)
                                     : umma_smem_desc_sw128(sA + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc_sw128(sB + k * 2048, BK * 128, 1024)
                                     : umma_smem_desc_sw128(sB + k * 32, 16, 1024);
            umma_bf16_ss(d_tmem, da, db, idesc, (g > t.g0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * stage);                    // smem slot free once MMAs retire
          if (g == t.g1 - 1) umma_commit(tfull0 + 8 * acc);   // accumulator ready for the epilogue
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // =============================== epilogue ===============================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int ew = warp - kEpiWarp0;
    // this warp's staging slice: the fp32 transpose staging of the staged epilogues, or the TMA-store tiles + bias slots
    const uint32_t stg = smem_u32(staging) + static_cast<uint32_t>(ew * kStagingFloats * 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    // one register-resident mode selects the epilogue instantiation (decided once, not per chunk)
    int mode = 0;  // 0 generic
    if (p.vec_ok) {
      const int res_mode = p.residual == nullptr ? 0 : (p.r_bf16 ? 2 : 1);
      const int e = p.epilogue == VITB_EPI_GELU ? 1 : (p.epilogue == VITB_EPI_GELU_BWD ? 2 : 3 + res_mode);
      mode = p.accumulate ? 1 : (p.d_bf16 ? 1 + e : 6 + e);
    }
    // measured (profiles/gemm_bench_r01.txt): the register-layout epilogue wins for plain / bias / GELU outputs,
    // but GELU' is faster with coalesced pre-activation loads in the staged layout
    const bool tma_path = p.tma_store != 0 && (mode == 2 || mode == 4);
    uint4 side_raw[8];                            // staged epilogue: side operand of the chunk in flight (see epi_vec_body)
    const uint32_t tbuf = (stg + 511u) & ~511u;   // NT 2 KiB 64B-swizzled tiles inside this warp's staging slice
    int tma_which = 0;
    if (tma_path && lane == 0) { tma_prefetch_desc(&tmD); if (p.D2 != nullptr) tma_prefetch_desc(&tmD2); }
    // packed GELU + GELU' path: the two 2 KiB store tiles leave 256 B of this warp's 4352 B slice unused (before the
    // tiles when the slice starts 256 B past a 512 B boundary, after them otherwise): two 32-float bias slots
    const bool packed = tma_path && mode == 2 && p.packed_epi != 0 && p.D2 != nullptr && p.d2_grad != 0;
    const uint32_t bias_slots = (tbuf == stg) ? stg + static_cast<uint32_t>(NT) * kTmaTileBytes : stg;
    const bool has_bias = p.bias != nullptr;
    const bool rowmul = p.rowmul != 0 && mode == 3;   // bf16 MUL_AUX without bias: register layout + TMA stores
    uint4 aux_next[4];                                  // this lane's 64 bytes of aux for the chunk that comes next
    if (rowmul && lane == 0) tma_prefetch_desc(&tmD);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int n0 = t.n_blk * BN;
      const int row_base = t.m_blk * BM + q * 32;
      // bias-type terms are added exactly once: by the split that owns the first k-block
      const bool lead_split = (t.g0 == 0);
      const int half = ew >> 2;  // which group of CHUNKS 32-column chunks of the tile this warp drains
      // grouped fp32 accumulate output: group g of the columns starts d_gs (not Ng) elements after group g - 1
      const long long d_off = (p.d_gs != 0 && p.group_cols > 0) ? static_cast<long long>(n0 / p.group_cols) * (p.d_gs - p.group_cols) : 0;
      if (rowmul) {
        aux_rows_load(p, lane, row_base, n0 + half * CHUNKS * 32, aux_next);
      } else {   // the side operand of this warp's first chunk is requested before the accumulator is waited for
        const int colf = n0 + half * CHUNKS * 32;
        if (colf < p.N) {
          switch (mode) {
            case 3: side_fetch<true, VITB_EPI_GELU_BWD, 0>(p, lane, row_base, colf, lead_split, side_raw); break;
            case 5: side_fetch<true, VITB_EPI_NONE, 1>(p, lane, row_base, colf, lead_split, side_raw); break;
            case 6: side_fetch<true, VITB_EPI_NONE, 2>(p, lane, row_base, colf, lead_split, side_raw); break;
            case 8: side_fetch<false, VITB_EPI_GELU_BWD, 0>(p, lane, row_base, colf, lead_split, side_raw); break;
            case 10: side_fetch<false, VITB_EPI_NONE, 1>(p, lane, row_base, colf, lead_split, side_raw); break;
            case 11: side_fetch<false, VITB_EPI_NONE, 2>(p, lane, row_base, colf, lead_split, side_raw); break;
            default: break;
          }
        }
      }
      float bias_next = 0.f;   // packed path: lane j carries bias[col0 + j] of the chunk that comes next
      if (packed && has_bias) {
        const int colb = n0 + half * CHUNKS * 32 + lane;
        if (colb < p.N) bias_next = __ldg(p.bias + colb);
      }
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = half * CHUNKS; c < (half + 1) * CHUNKS; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;
        const int next_col0 = (c + 1 < (half + 1) * CHUNKS && col0 + 32 < p.N) ? col0 + 32 : -1;
        uint32_t r[32];
#ifdef VITB_GEMM_DIAG
        if (VITB_DIAG(32)) continue;
        if (VITB_DIAG(4)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        } else
#endif
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * BN + c * 32), r);
        if (packed) {
          const uint32_t slot = bias_slots + static_cast<uint32_t>(c & 1) * 128u;
          if (has_bias) {
            // publish this chunk's bias (requested a chunk ago) and request the next chunk's: the L2 round trip
            // runs under this chunk's math
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot + static_cast<uint32_t>(lane) * 4u), "f"(bias_next) : "memory");
            bias_next = 0.f;
            if (next_col0 >= 0 && next_col0 + lane < p.N) bias_next = __ldg(p.bias + next_col0 + lane);
            __syncwarp();
          }
          tmem_ld_wait();
          epi_rows_gelu_dg_packed<NT>(&tmD, &tmD2, tbuf, tma_which, lane, row_base, col0, r, slot, has_bias);
          continue;
        }
        if (rowmul) {
          uint4 aux_cur[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) aux_cur[u] = aux_next[u];
          if (next_col0 >= 0) aux_rows_load(p, lane, row_base, next_col0, aux_next);   // flies during this chunk
          tmem_ld_wait();
          epi_rows_mul_aux<NT>(p, &tmD, tbuf, tma_which, lane, row_base, col0, r, aux_cur);
          continue;
        }
        tmem_ld_wait();
        if (tma_path) {    // bf16 outputs without residual / column sums: math in registers, tiles leave by TMA
          if (mode == 2) epi_rows_bf16_tma<VITB_EPI_GELU, NT>(p, &tmD, &tmD2, tbuf, tma_which, lane, row_base, col0, r);
          else epi_rows_bf16_tma<VITB_EPI_NONE, NT>(p, &tmD, &tmD2, tbuf, tma_which, lane, row_base, col0, r);
          continue;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) st_shared_v2(stg + lane * (kStgStride * 4) + j * 8, r[2 * j], r[2 * j + 1]);
        __syncwarp();
        switch (mode) {
          case 1: epi_vec<false, VITB_EPI_NONE, 0, true>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0, d_off); break;
          case 2: epi_vec<true, VITB_EPI_GELU, 0, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 3: epi_vec<true, VITB_EPI_GELU_BWD, 0, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 4: epi_vec<true, VITB_EPI_NONE, 0, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 5: epi_vec<true, VITB_EPI_NONE, 1, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 6: epi_vec<true, VITB_EPI_NONE, 2, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 7: epi_vec<false, VITB_EPI_GELU, 0, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 8: epi_vec<false, VITB_EPI_GELU_BWD, 0, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 9: epi_vec<false, VITB_EPI_NONE, 0, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 10: epi_vec<false, VITB_EPI_NONE, 1, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          case 11: epi_vec<false, VITB_EPI_NONE, 2, false>(p, stg, lane, row_base, col0, lead_split, side_raw, next_col0); break;
          default: epi_generic(p, stg, lane, row_base, col0, lead_split); break;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if ((tma_path || rowmul) && lane == 0) bulk_wait_all();   // staging tiles must outlive the stores that read them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ================================================================================================
// Weight-gradient GEMM on CTA pairs (tcgen05 cta_group::2): dW[M,N] += A[K,M]^T B[K,N], both operands
// MN-major, K = tokens, fp32 red.global accumulation, split along K.
//
// The weight-gradient shapes are the ones a pair tiles exactly (M, N in {768, 3072}: multiples of 256) and
// the ones furthest from cuBLAS in round 1 (1.10-1.16 vs 1.33-1.37 PFLOP/s, profiles/gemm_bench_r01.txt):
// with K = 25,216 the mainloop is everything, and a 256 x 256 pair tile fetches 128 + 128 operand rows per SM
// and k-block instead of 128 + 256 — a third less L2 -> SM traffic and shared-memory fill per FLOP.
// (The token-major GEMMs stay on the single-CTA kernel: T = 197 x 128 rows leave half a pair tile over, which
// costs a whole extra wave at N = 768.)
//
// Protocol (per k-block stage s, 6 stages of 32 KiB per CTA):
//   both CTAs : producer lane waits its OWN empty[s], then TMA-loads its A and B halves with
//               .cta_group::2 completion on the LEADER's full[s]; the leader alone arms expect_tx (2 x stage)
//   leader    : waits full[s], issues 4 x tcgen05.mma.cta_group::2 (256 x 256 x 16), commits with
//               .multicast::cluster to empty[s] of BOTH CTAs; after the last k-block to tmem_full[acc] of both
//   both CTAs : epilogue warps drain their own 128 TMEM lanes (rows 128*rank ..) and arrive on the LEADER's
//               tmem_empty[acc] (count = 2 x 8 warps)
// ================================================================================================
constexpr int kPairStages = 6;
constexpr int kPairStageBytes = 2 * A_BYTES;   // A half 128 x 64 + B half 128 x 64, bf16
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kStagingBytes + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
vitb_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ GemmDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("vitb_wgrad_pair: dynamic shared memory base 0x%x is not 1024-byte aligned\n", smem_u32(smem));
    __trap();
  }
  float* staging = reinterpret_cast<float*>(smem + kPairStages * kPairStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPairStages * kPairStageBytes + kStagingBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = smem_u32(bars + kPairStages);
  const uint32_t tfull0 = smem_u32(bars + 2 * kPairStages);
  const uint32_t tempty0 = smem_u32(bars + 2 * kPairStages + 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, 2 * kEpiWarps);   // the epilogue warps of BOTH CTAs arrive on the leader's copy
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(tmem_slot), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();     // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.split_k;     // pair tiles (256 x 256 x K-split)
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const TileCoord t = decode_tile(p, tile);
        const int m0 = t.m_blk * 256 + static_cast<int>(rank) * 128;
        const int n0 = t.n_blk * 256 + static_cast<int>(rank) * 128;
        for (int g = t.g0; g < t.g1; ++g) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1u);
          if (leader) mbar_arrive_expect_tx(full0 + 8 * stage, 2 * kPairStageBytes);
          const uint32_t fb = mapa_shared(full0 + 8 * stage, 0);   // the leader's barrier collects both halves
          const uint32_t sA = smem_u32(smem + stage * kPairStageBytes);
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int i = 0; i < 2; ++i) tma_load_2d_pair(&tmA, fb, sA + i * 8192, m0 + 64 * i, g * BK);
#pragma unroll
          for (int i = 0; i < 2; ++i) tma_load_2d_pair(&tmB, fb, sB + i * 8192, n0 + 64 * i, g * BK);
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256, true, true);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const TileCoord t = decode_tile(p, tile);
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int g = t.g0; g < t.g1; ++g) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * kPairStageBytes);
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_ss_pair(d_tmem, umma_smem_desc_sw128(sA + k * 2048, BK * 128, 1024),
                              umma_smem_desc_sw128(sB + k * 2048, BK * 128, 1024), idesc, (g > t.g0 || k > 0) ? 1u : 0u);
          umma_commit_pair(empty0 + 8 * stage, 3);                    // both CTAs' smem slots free once MMAs retire
          if (g == t.g1 - 1) umma_commit_pair(tfull0 + 8 * acc, 3);   // both CTAs' accumulator halves ready
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // =============================== epilogue (both CTAs): fp32 red.global accumulation ===============================
    const int q = warp & 3;
    const uint32_t stg = smem_u32(staging) + static_cast<uint32_t>((warp - kEpiWarp0) * kStagingFloats * 4);
    const uint32_t tempty_leader = mapa_shared(tempty0, 0);
    uint4 no_side[8];   // accumulate mode has no side operand (never read)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      const TileCoord t = decode_tile(p, tile);
      const int n0 = t.n_blk * 256;
      const int row_base = t.m_blk * 256 + static_cast<int>(rank) * 128 + q * 32;
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      const bool lead_split = (t.g0 == 0);
      const int half = (warp - kEpiWarp0) >> 2;
      const long long d_off = (p.d_gs != 0 && p.group_cols > 0) ? static_cast<long long>(n0 / p.group_cols) * (p.d_gs - p.group_cols) : 0;
#pragma unroll 1
      for (int c = half * 4; c < (half + 1) * 4; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 256 + c * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) st_shared_v2(stg + lane * (kStgStride * 4) + j * 8, r[2 * j], r[2 * j + 1]);
        __syncwarp();
        epi_vec<false, VITB_EPI_NONE, 0, true>(p, stg, lane, row_base, col0, lead_split, no_side, -1, d_off);
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  cluster_sync_all();     // the peer's shared memory and TMEM are in use until the leader's last MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

int launch_wgrad_pair(const CUtensorMap* tm, GemmDev d, int M, int N, cudaStream_t stream) {
  const int sms = vitb_num_sms();
  const int npairs = sms / 2;
  d.m_tiles = (M + 255) / 256;
  d.n_tiles = (N + 255) / 256;
  const int tiles = d.m_tiles * d.n_tiles;
  int split = npairs / (tiles > 0 ? tiles : 1);
  if (split > d.total_kblocks / 4) split = d.total_kblocks / 4;
  if (split < 1) split = 1;
  d.split_k = split;
  const long long total = (long long)tiles * split;
  const int grid = 2 * (int)(total < npairs ? total : npairs);
  VITB_CUDA_CHECK(cudaFuncSetAttribute(vitb_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
  vitb_wgrad_pair_kernel<<<grid, kThreads, kPairSmemBytes, stream>>>(tm[0], tm[1], d);
  VITB_LAUNCH_CHECK("vitb_wgrad_pair_kernel");
  return VITB_OK;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const CUtensorMap* tm, const GemmDev& d, int grid, cudaStream_t stream) {
  auto kern = vitb_gemm_kernel<BN, A_MN, B_MN>;
  VITB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg<BN>::SMEM_BYTES));
  VITB_CUDA_CHECK(vitb_launch<kPdlGemm>(kern, dim3(grid), dim3(kThreads), Cfg<BN>::SMEM_BYTES, stream, tm[0], tm[1], tm[2], tm[3],
                              tm[4], tm[5], tm[6], tm[7], d));
  VITB_LAUNCH_CHECK("vitb_gemm_kernel");
  return VITB_OK;
}

}  // namespace

#ifdef VITB_GEMM_DIAG
extern "C" int vitb_gemm_diag(const vitb_gemm_params* p, void* stream);
extern "C" int vitb_gemm_diag_mask(int mask) {   // synchronising; ablation tool only
  VITB_CUDA_CHECK(cudaMemcpyToSymbol(g_diag_mask, &mask, sizeof(int)));
  return VITB_OK;
}
#endif

extern "C" int VITB_GEMM_ENTRY(const vitb_gemm_params* p, void* stream_) {
  VITB_REQUIRE(p != nullptr, VITB_ERR_BAD_ARG, "vitb_gemm: null params");
  VITB_REQUIRE(p->struct_bytes == (int)sizeof(vitb_gemm_params), VITB_ERR_BAD_ARG,
               "vitb_gemm: struct_bytes %d != %d (ABI mismatch)", p->struct_bytes,
               (int)sizeof(vitb_gemm_params));
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VITB_REQUIRE(p->M >= 0 && p->N >= 0, VITB_ERR_BAD_ARG, "vitb_gemm: negative M/N");
  if (p->M == 0 || p->N == 0) return VITB_OK;  // empty problem: nothing to launch
  VITB_REQUIRE(p->num_segments >= 1 && p->num_segments <= 3, VITB_ERR_BAD_ARG,
               "vitb_gemm: num_segments %d", p->num_segments);
  VITB_REQUIRE(p->D != nullptr, VITB_ERR_BAD_ARG, "vitb_gemm: D is null");
  VITB_REQUIRE(p->d_dtype == VITB_F32 || p->d_dtype == VITB_BF16, VITB_ERR_BAD_ARG, "vitb_gemm: d_dtype");
  VITB_REQUIRE(!(p->accumulate && p->d_dtype != VITB_F32), VITB_ERR_BAD_ARG,
               "vitb_gemm: accumulate needs an fp32 D");
  VITB_REQUIRE(!(p->split_k > 1 && !p->accumulate), VITB_ERR_BAD_ARG,
               "vitb_gemm: split_k > 1 needs accumulate = 1");
  VITB_REQUIRE(p->epilogue >= 0 && p->epilogue <= 4, VITB_ERR_BAD_ARG, "vitb_gemm: epilogue %d", p->epilogue);
  VITB_REQUIRE(!((p->epilogue == VITB_EPI_GELU_BWD || p->epilogue == VITB_EPI_MUL_AUX) && p->aux == nullptr),
               VITB_ERR_BAD_ARG, "vitb_gemm: GELU_BWD / MUL_AUX need aux");
  VITB_REQUIRE(!(p->epilogue == VITB_EPI_GELU_DG && p->D2 == nullptr), VITB_ERR_BAD_ARG, "vitb_gemm: GELU_DG needs D2");
  VITB_REQUIRE(!(p->accumulate && p->epilogue != VITB_EPI_NONE && p->split_k != 1), VITB_ERR_BAD_ARG,
               "vitb_gemm: a non-linear epilogue cannot be split along K");
  VITB_REQUIRE(p->row_bias == nullptr || p->row_bias_group > 0, VITB_ERR_BAD_ARG,
               "vitb_gemm: row_bias_group must be > 0");
  const int ngroups = p->n_groups > 1 ? p->n_groups : 1;
  VITB_REQUIRE(p->N % ngroups == 0, VITB_ERR_BAD_ARG, "vitb_gemm: N %d is not a multiple of n_groups %d", p->N, ngroups);
  VITB_REQUIRE(ngroups == 1 || p->b_group_stride % 8 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "vitb_gemm: b_group_stride %% 8");
  VITB_REQUIRE(p->d_group_stride == 0 || (ngroups > 1 && p->accumulate && p->d_dtype == VITB_F32 && p->epilogue == VITB_EPI_NONE &&
                                          p->d_group_stride % 4 == 0),
               VITB_ERR_BAD_ARG, "vitb_gemm: a grouped output needs n_groups > 1, accumulate = 1, fp32 D, no epilogue");

  GemmDev d{};
  d.M = p->M;
  d.N = p->N;
  d.nseg = p->num_segments;
  d.total_kblocks = 0;
  for (int s = 0; s < 3; ++s) d.kblocks[s] = 0;
  for (int s = 0; s < p->num_segments; ++s) {
    VITB_REQUIRE(p->K[s] > 0, VITB_ERR_BAD_ARG, "vitb_gemm: K[%d] = %d", s, p->K[s]);
    VITB_REQUIRE(p->A[s] != nullptr && p->B[s] != nullptr, VITB_ERR_BAD_ARG, "vitb_gemm: null operand %d", s);
    d.kblocks[s] = (p->K[s] + BK - 1) / BK;
    d.total_kblocks += d.kblocks[s];
  }
  // tile width: 256 unless that pads N more than 128 would
  const int pad256 = ((p->N + 255) / 256) * 256 - p->N;
  const int pad128 = ((p->N + 127) / 128) * 128 - p->N;
  // (128-column tiles for the narrow N = 768 outputs, whose 3-4 wide tiles per CTA leave the first mainloop and the last
  // epilogue uncovered, were measured: no gain for the out-projection, 30-50 % slower elsewhere, profiles/gemm_narrow_tiles_r02.txt)
  const int BN = (pad256 <= pad128) ? 256 : 128;
  const int Ng = p->N / ngroups;
  VITB_REQUIRE(ngroups == 1 || Ng % BN == 0, VITB_ERR_UNSUPPORTED_SHAPE,
               "vitb_gemm: group width %d is not a multiple of the %d-column tile", Ng, BN);
  VITB_REQUIRE(p->m_dev == nullptr || (!p->a_mn_major && p->split_k <= 1 && p->colsum == nullptr && !p->accumulate),
               VITB_ERR_BAD_ARG, "vitb_gemm: m_dev needs a token-major, non-accumulating GEMM without split-K / colsum");
  d.m_dev = p->m_dev;
  d.group_cols = ngroups > 1 ? Ng : 0;
  d.d_gs = ngroups > 1 ? p->d_group_stride : 0;
  d.m_tiles = (p->M + BM - 1) / BM;
  d.n_tiles = (p->N + BN - 1) / BN;
  const int sms = vitb_num_sms();
  int split = p->split_k;
  if (split <= 0) {
    split = 1;
    if (p->accumulate && p->epilogue == VITB_EPI_NONE) {
      const int tiles = d.m_tiles * d.n_tiles;
      split = sms / (tiles > 0 ? tiles : 1);
      if (split < 1) split = 1;
      // keep at least 4 k-blocks per split so the pipeline fill is amortised
      if (split > d.total_kblocks / 4) split = d.total_kblocks / 4;
      if (split < 1) split = 1;
    }
  }
  VITB_REQUIRE(split <= d.total_kblocks, VITB_ERR_BAD_ARG, "vitb_gemm: split_k %d > k-blocks %d", split,
               d.total_kblocks);
  d.split_k = split;
  // the two "derivative carried by the forward" flavours reuse the GELU / GELU_BWD code paths
  d.d2_grad = p->epilogue == VITB_EPI_GELU_DG;
  d.aux_grad = p->epilogue == VITB_EPI_MUL_AUX;
  d.epilogue = d.d2_grad ? VITB_EPI_GELU : (d.aux_grad ? VITB_EPI_GELU_BWD : p->epilogue);
  d.D = p->D;
  d.ldd = p->ldd;
  d.d_bf16 = p->d_dtype == VITB_BF16;
  d.accumulate = p->accumulate;
  d.D2 = p->D2;
  d.ldd2 = p->ldd2;
  d.bias = p->bias;
  d.row_bias = p->row_bias;
  d.row_bias_group = p->row_bias_group;
  d.row_remap_group = p->row_remap_group;
  d.residual = p->residual;
  d.ldr = p->ldr;
  d.r_bf16 = p->r_dtype == VITB_BF16;
  d.aux = p->aux;
  d.ldaux = p->ldaux;
  {
    auto al = [](const void* q, uintptr_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
    const uintptr_t d_al = d.d_bf16 ? 8 : 16;
    d.vec_ok = (p->N % 4 == 0) && (p->ldd % 4 == 0) && al(p->D, d_al) && p->row_bias == nullptr &&
               p->row_remap_group == 0 && al(p->bias, 16) &&
               (p->residual == nullptr || (p->ldr % 4 == 0 && al(p->residual, d.r_bf16 ? 8 : 16))) &&
               (p->D2 == nullptr || (p->ldd2 % 4 == 0 && al(p->D2, d_al))) &&
               (p->aux == nullptr || (p->ldaux % 4 == 0 && al(p->aux, d_al))) &&
               !(p->accumulate && (p->residual != nullptr || d.epilogue != VITB_EPI_NONE)) &&
               !(d.epilogue != VITB_EPI_NONE && p->residual != nullptr);
    d.colsum = p->colsum;
    VITB_REQUIRE(d.d_gs == 0 || d.vec_ok, VITB_ERR_UNSUPPORTED_SHAPE,
                 "vitb_gemm: a grouped output needs the vectorised epilogue (N, ldd multiples of 4, 16-byte aligned D)");
    VITB_REQUIRE(p->colsum == nullptr || (d.vec_ok && al(p->colsum, 16) && !p->accumulate), VITB_ERR_UNSUPPORTED_SHAPE,
                 "vitb_gemm: colsum needs the vectorised epilogue (N, lds multiples of 4, 16-byte aligned pointers)");
  }

  CUtensorMap tm[8];
  for (int s = 0; s < 3; ++s) {
    const int src = s < p->num_segments ? s : 0;
    const uint64_t K = (uint64_t)p->K[src];
    if (!p->a_mn_major)
      st = vitb_make_tmap_2d_bf16(&tm[2 * s], p->A[src], K, (uint64_t)p->M, (uint64_t)p->lda[src] * 2, BK, BM);
    else
      st = vitb_make_tmap_2d_bf16(&tm[2 * s], p->A[src], (uint64_t)p->M, K, (uint64_t)p->lda[src] * 2, 64, BK);
    if (st != VITB_OK) return st;
    {
      // 3-D: (inner, outer, column group).  One group: the group stride is never used (any 16-byte multiple will do).
      const uint64_t gs_bytes = ngroups > 1 ? (uint64_t)p->b_group_stride * 2 : 16;
      uint64_t dims[3], str[2] = {(uint64_t)p->ldb[src] * 2, gs_bytes};
      uint32_t box[3];
      if (!p->b_mn_major) { dims[0] = K; dims[1] = (uint64_t)Ng; box[0] = BK; box[1] = (uint32_t)BN; }
      else { dims[0] = (uint64_t)Ng; dims[1] = K; box[0] = 64; box[1] = BK; }
      dims[2] = (uint64_t)ngroups;
      box[2] = 1;
      st = vitb_make_tmap_nd_bf16(&tm[2 * s + 1], p->B[src], 3, dims, str, box);
    }
    if (st != VITB_OK) return st;
  }
  // bf16 outputs of the register-layout epilogue leave through TMA stores (32 x 32 tiles, SWIZZLE_64B);
  // anything that does not meet TMA's 16-byte rules takes the staged epi_vec path
  d.tma_store = 0;
  {
    // packed-pair GELU + GELU' epilogue: on (measured, profiles/epi_ab_r01d.txt: fc1 forward 0.145 -> 0.133 ms);
    // VITB_EPI_PACKED=0 selects the scalar epilogue for A/B runs
    const char* pe = getenv("VITB_EPI_PACKED");
    d.packed_epi = (pe == nullptr || atoi(pe) != 0) ? 1 : 0;
  }
  tm[6] = tm[0];
  tm[7] = tm[0];
  if (d.vec_ok && d.d_bf16 && p->N % 8 == 0 && p->colsum == nullptr && p->residual == nullptr && !p->accumulate &&
      d.epilogue != VITB_EPI_GELU_BWD && p->ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(p->D) & 15u) == 0 &&
      (p->D2 == nullptr || (p->ldd2 % 8 == 0 && (reinterpret_cast<uintptr_t>(p->D2) & 15u) == 0))) {
    st = vitb_make_tmap_2d_bf16_sw64(&tm[6], p->D, (uint64_t)p->N, (uint64_t)p->M, (uint64_t)p->ldd * 2, 32, 32);
    if (st != VITB_OK) return st;
    if (p->D2 != nullptr) {
      st = vitb_make_tmap_2d_bf16_sw64(&tm[7], p->D2, (uint64_t)p->N, (uint64_t)p->M, (uint64_t)p->ldd2 * 2, 32, 32);
      if (st != VITB_OK) return st;
    }
    d.tma_store = 1;
  }
  d.rowmul = 0;
  {
    // register-layout MUL_AUX epilogue: on (measured, profiles/epi_ab_r01d.txt: fc2 dgrad 0.144 -> 0.131 ms);
    // VITB_EPI_ROWMUL=0 selects the staged epilogue for A/B runs
    const char* re = getenv("VITB_EPI_ROWMUL");
    const bool want = re == nullptr || atoi(re) != 0;
    if (want && d.aux_grad && d.vec_ok && d.d_bf16 && p->N % 8 == 0 && p->residual == nullptr && !p->accumulate &&
        p->bias == nullptr && p->D2 == nullptr && p->ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(p->D) & 15u) == 0 &&
        p->ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(p->aux) & 15u) == 0 &&
        (p->colsum == nullptr || (reinterpret_cast<uintptr_t>(p->colsum) & 3u) == 0)) {
      st = vitb_make_tmap_2d_bf16_sw64(&tm[6], p->D, (uint64_t)p->N, (uint64_t)p->M, (uint64_t)p->ldd * 2, 32, 32);
      if (st != VITB_OK) return st;
      d.rowmul = 1;
    }
  }
  // weight gradients (both operands MN-major, fp32 accumulation, one K segment) run on CTA pairs
  // (measured, profiles/gemm_bench_r01b.txt: dW fc1 0.099 -> 0.090 ms, dW fc2 0.104 -> 0.089 ms, i.e. 1.33 PFLOP/s)
  {
    const char* pair_env = getenv("VITB_GEMM_PAIR");   // VITB_GEMM_PAIR=0 keeps the single-CTA kernel (A/B measurements)
    const bool pair_on = pair_env == nullptr || atoi(pair_env) != 0;
    if (pair_on && p->a_mn_major && p->b_mn_major && p->accumulate && p->num_segments == 1 &&
        d.epilogue == VITB_EPI_NONE && !d.d_bf16 && d.vec_ok && p->bias == nullptr && p->colsum == nullptr &&
        p->split_k <= 0 && p->M >= 256 && p->N >= 256 && sms >= 2 && (ngroups == 1 || (p->d_group_stride != 0 && Ng % 256 == 0))) {
      // the pair kernel reads B through a plain 2-D map: a grouped OUTPUT is fine, a grouped B is not needed here
      // (the merged weight gradient contracts against the packed [T, 3 Ng] dqkv buffer)
      VITB_REQUIRE(ngroups == 1 || p->b_group_stride == Ng, VITB_ERR_UNSUPPORTED_SHAPE,
                   "vitb_gemm: the weight-gradient pair kernel needs a column-contiguous B (b_group_stride == N / n_groups)");
      st = vitb_make_tmap_2d_bf16(&tm[1], p->B[0], (uint64_t)p->N, (uint64_t)p->K[0], (uint64_t)p->ldb[0] * 2, 64, BK);
      if (st != VITB_OK) return st;
      return launch_wgrad_pair(tm, d, p->M, p->N, stream);
    }
  }
  const long long total_tiles = (long long)d.m_tiles * d.n_tiles * d.split_k;
  const int grid = (int)(total_tiles < sms ? total_tiles : sms);

#define VITB_DISPATCH(BN_)                                                              \
  do {                                                                                  \
    if (!p->a_mn_major && !p->b_mn_major) return launch<BN_, false, false>(tm, d, grid, stream); \
    if (!p->a_mn_major && p->b_mn_major) return launch<BN_, false, true>(tm, d, grid, stream);   \
    if (p->a_mn_major && !p->b_mn_major) return launch<BN_, true, false>(tm, d, grid, stream);   \
    return launch<BN_, true, true>(tm, d, grid, stream);                                \
  } while (0)
  if (BN == 256) VITB_DISPATCH(256);
  VITB_DISPATCH(128);
#undef VITB_DISPATCH
}
