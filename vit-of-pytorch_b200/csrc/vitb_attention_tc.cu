// vitb_attention_tc.cu — flash-style SelfAttention on tcgen05 tensor cores (bf16 in, fp32 in TMEM).
//
// Replaces the q k^T / softmax / v chain of SelfAttention.forward (src/model.py:90-97) and
// Attention.forward (res-vit/model.py:273-293) plus its autograd backward, for head_dim 64 and up
// to 256 keys (ViT-B/L at 224 px: N = 197 or 50).  The [B,H,N,N] score tensor the reference
// materialises in HBM never leaves the SM: S lives in TMEM, P goes through shared memory straight
// into the second MMA.  The non-power-of-two tail (197 = 128 + 69 query rows, 208 = 13 x 16 key
// columns) is handled by TMA zero fill plus explicit masking of key columns >= N.
//
// Forward, one CTA per (image, head, 128-query tile), 2 CTAs resident per SM:
//   TMA   : Q[128x64], K[NKx64], V[NKx64] straight from the packed [T, 3D] projection output
//   MMA 1 : S = Q K^T        (128 x NK x 64, one accumulator of NK <= 256 TMEM columns)
//   warps : row max / exp2 / row sum in fp32 (one thread per query row), P -> bf16 -> swizzled smem
//   MMA 2 : O = P V          (128 x 64 x NK, V consumed MN-major from the same TMA image)
//   warps : O / rowsum -> bf16 -> HBM, LSE -> HBM
// Because all keys fit one accumulator there is no online-softmax rescale pass at these sizes.
//
// Backward, one CTA per (image, head); every Q / dO / O tile of the head is fetched by TMA up front, the CTA
// loops over the query tiles, dK/dV accumulate in TMEM:
//   D = rowsum(dO * O) (from the smem tiles) ; S = Q K^T ; P = exp(S*c - LSE) ; dP = dO V^T ;
//   dS = P * (dP - D) * c (in place over P) ; dV += P^T dO ; dK += dS^T Q ; dQ = dS K   (5 tcgen05 GEMMs per tile)
#include "../../include/vitb200.h"
#include <stdlib.h>

#include "vitb_common.cuh"
#include "vitb_attn_util.cuh"

namespace {
using namespace vitb;
using namespace vitb::attn;

struct AttnTc {
  int N;    // tokens (queries == keys)
  int NK;   // keys padded to a multiple of 16
  int H;
  float scale;       // 1/sqrt(dh)
  float scale_log2;  // scale * log2(e)
  float* lse;  // [B,H,N]   (q, k, v, o, dO and the gradients move through tensor maps)
  // attn_bwd_tc<true> (more than 256 tokens): dQ of a key block is ADDED into this fp32 [B, N, H*64] buffer
  float* dq_acc;
};

// ================================================================================================
// forward
// ================================================================================================
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
            const __grid_constant__ AttnTc a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int NK = a.NK;
  const int kv_bytes = NK * 128;
  const int nchunks = (NK + 63) >> 6;
  const int u_bytes = max(kChunkBytes + kv_bytes, nchunks * kChunkBytes);
  uint8_t* sV = smem;
  uint8_t* sU = smem + kv_bytes;       // Q | K, later overwritten by P
  uint8_t* sQ = sU;
  uint8_t* sK = sU + kChunkBytes;
  uint8_t* sP = sU;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sU + u_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* red = reinterpret_cast<float*>(bars + 8);   // [2 stats][2 halves][128 rows]
  const uint32_t bar_qk = smem_u32(bars), bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_o = bar_qk + 24;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int row0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  pdl_trigger();

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // first global access (the TMA loads) comes after the previous grid has completed

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_qk, kChunkBytes + kv_bytes);
    tma_load_3d(&tmQ, bar_qk, smem_u32(sQ), h * DH, row0, b);
    tma_load_3d(&tmK, bar_qk, smem_u32(sK), h * DH, 0, b);
    mbar_arrive_expect_tx(bar_v, kv_bytes);
    tma_load_3d(&tmV, bar_v, smem_u32(sV), h * DH, 0, b);
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, NK, false, false);
#pragma unroll
    for (int k = 0; k < DH / 16; ++k)
      umma_bf16_ss(tmem_base, umma_smem_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024),
                   umma_smem_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
    umma_commit(bar_s);
  }
  __syncwarp();
  mbar_wait(bar_s, 0);
  tc_fence_after();

  const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
  const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const ColRange cr = my_chunks(NK, half);
  // pass 1: row max over the valid keys of this thread's column half (two TMEM loads in flight per wait)
  float mx = -INFINITY;
  for (int c = cr.c_begin; c < cr.c_end; c += 2) {
    uint32_t v0[32], v1[32];
    const bool two = (c + 1 < cr.c_end);
    issue_chunk(trow, c * 32, NK, v0);
    if (two) issue_chunk(trow, (c + 1) * 32, NK, v1);
    tmem_ld_wait();
    mx = fmaxf(mx, chunk_max(v0, c * 32, NK, a.N));
    if (two) mx = fmaxf(mx, chunk_max(v1, (c + 1) * 32, NK, a.N));
  }
  red[half * 128 + r] = mx;
  __syncthreads();
  mx = fmaxf(red[r], red[128 + r]);
  // pass 2: p = exp2((s - max) * c), row sum, bf16 P into the swizzled A-operand image
  float sum = 0.f;
  const float mxs = mx * a.scale_log2;
  const uint32_t sP_u = smem_u32(sP);
  for (int c = cr.c_begin; c < cr.c_end; ++c) {
    const int c0 = c * 32;
    uint32_t v[32];
    ld_chunk(trow, c0, NK, 0u, v);
    float pv[32];
    if (c0 + 32 <= a.N) {   // interior chunk: every column is a valid key
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        pv[j] = ex2_approx(fmaf(__uint_as_float(v[j]), a.scale_log2, -mxs));
        sum += pv[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = ex2_approx(fmaf(__uint_as_float(v[j]), a.scale_log2, -mxs));
        pv[j] = (c0 + j < a.N) ? e : 0.f;
        sum += pv[j];
      }
    }
    const int kc = c0 >> 6, u0 = (c0 & 63) >> 3, nunits = (c0 + 32 <= NK) ? 4 : 2;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < nunits)
        st_shared_v4(sP_u + kc * kChunkBytes + swz_unit(r, u0 + u), pack_bf16x2(pv[8 * u + 0], pv[8 * u + 1]),
                     pack_bf16x2(pv[8 * u + 2], pv[8 * u + 3]), pack_bf16x2(pv[8 * u + 4], pv[8 * u + 5]),
                     pack_bf16x2(pv[8 * u + 6], pv[8 * u + 7]));
    }
  }
  red[256 + half * 128 + r] = sum;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  sum = red[256 + r] + red[256 + 128 + r];
  if (tid == 0) {
    tc_fence_after();
    mbar_wait(bar_v, 0);
    const uint32_t idesc = umma_idesc_bf16(128, DH, false, true);
    const int nks = NK >> 4;
    for (int t = 0; t < nks; ++t)
      umma_bf16_ss(tmem_base, umma_smem_desc_sw128(sP_u + (t >> 2) * kChunkBytes + (t & 3) * 32, 16, 1024),
                   umma_smem_desc_sw128(smem_u32(sV) + t * 2048, 8192, 1024), idesc, t > 0 ? 1u : 0u);
    umma_commit(bar_o);
  }
  __syncwarp();
  mbar_wait(bar_o, 0);
  tc_fence_after();
  const int row = row0 + r;
  {
    uint32_t v[32];
    tmem_ld_32x32b_x32(trow + half * 32, v);   // this thread's 32 of the 64 head-dim columns
    tmem_ld_wait();
    stage_row32_bf16(sP_u, r, half, v, 1.0f / sum);   // the P image is dead: the PV MMA has retired
    if (row < a.N && a.lse && half == 0) a.lse[(static_cast<long long>(b) * a.H + h) * a.N + row] = mx * a.scale + logf(sum);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {   // O tile [128 x 64] leaves as one TMA store (rows >= N clipped)
    tma_store_3d(&tmO, sP_u, h * DH, row0, b);
    bulk_commit();
    bulk_wait_all();
  }
  if (warp == 0) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ================================================================================================
// forward, general shapes: 64 <= head_dim <= 128 (multiple of 16) and ANY number of keys — ViT-H/14 is
// head_dim 80 x 257 tokens at 224 px (src/config.py:93-104); the reference's default evaluation resolution of
// 384 px (src/config.py:12) gives 577 tokens for */16 and 730 for h14.  Same plan as attn_fwd_tc, plus:
//   * the head dimension spans two 64-column 128B-swizzled chunks; the second TMA box reads 64 columns even
//     when head_dim < 128 (the surplus belongs to the next head or is zero-filled past the row), but the
//     MMAs only ever consume head_dim columns: K-extent head_dim/16 steps for S, N = head_dim for O
//   * keys are processed in blocks of up to 320 (one block for ViT-H/14 at 224 px).  More than 256 keys do
//     not fit one MMA: S is issued as N = 256 plus the rest into adjacent TMEM columns, still one
//     accumulator row per query
//   * with several key blocks the running row max / row sum follow the online-softmax recurrence and O is
//     accumulated in registers: O <- O * exp2((m_old - m_new) c) + P_blk V_blk, each block's product read
//     back from TMEM (no TMEM read-modify-write)
// TMEM: S in columns [0, 320), O_blk in [384, 384 + head_dim); one CTA per SM.
// ================================================================================================
struct AttnGen {
  int N, H, dh;
  int KB;                  // keys per block (multiple of 16, <= 320); smem chunk stride = KB * 128 bytes
  int nblocks;
  int kv_loads, kv_rows;   // K / V chunks of a block arrive as kv_loads boxes of kv_rows rows (boxes hold <= 256 rows)
  float scale, scale_log2;
  __nv_bfloat16* o;
  long long o_bs, o_rs;
  float* lse;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_tc_gen(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnGen a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int dh = a.dh;
  const int kvc = a.KB * 128;                     // one 64-column chunk of a K or V block: [KB rows x 128 B]
  const int npchunks = (a.KB + 63) >> 6;          // 64-key chunks of the P image
  const int kp_bytes = max(2 * kvc, npchunks * kChunkBytes);
  uint8_t* sV = smem;                             // [2 chunks][KB x 64]
  uint8_t* sQ = sV + 2 * kvc;                     // [2 chunks][128 x 64], kept for every key block
  uint8_t* sK = sQ + 2 * kChunkBytes;             // [2 chunks][KB x 64], overwritten by P once S is in TMEM
  uint64_t* bars = reinterpret_cast<uint64_t*>(sK + kp_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* red = reinterpret_cast<float*>(bars + 8);   // [2 stats][2 halves][128 rows]
  const uint32_t bar_qk = smem_u32(bars), bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_o = bar_qk + 24;
  const uint32_t sV_u = smem_u32(sV), sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sP_u = sK_u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int row0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  pdl_trigger();

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 384u;
  pdl_wait();

  const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
  const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const int ksteps = dh >> 4;                     // head-dim steps of 16 for S = Q K^T
  // 16-column units of the head dimension owned by this thread: half 0 the first ceil(n/2), half 1 the rest
  const int nun = dh >> 4, umid = (nun + 1) >> 1;
  const int ub = half ? umid : 0, ue = half ? nun : umid;
  const bool multi = a.nblocks > 1;
  float o_acc[4][16];                             // running O (only used with several key blocks)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) o_acc[i][j] = 0.f;
  float m_run = -INFINITY;                        // running row max (identical in both halves of a row)
  float sum = 0.f;                                // this half's running row sum

  for (int kb = 0; kb < a.nblocks; ++kb) {
    const uint32_t ph = kb & 1;
    const int key0 = kb * a.KB;
    const int valid = min(a.N - key0, a.KB);      // keys of this block that exist
    const int NK = (valid + 15) & ~15;            // MMA / softmax extent of this block
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_qk, (kb == 0 ? 2 * kChunkBytes : 0) + 2 * kvc);
      for (int c = 0; c < 2; ++c) {
        if (kb == 0) tma_load_3d(&tmQ, bar_qk, sQ_u + c * kChunkBytes, h * dh + 64 * c, row0, b);
        for (int l = 0; l < a.kv_loads; ++l)
          tma_load_3d(&tmK, bar_qk, sK_u + c * kvc + l * a.kv_rows * 128, h * dh + 64 * c, key0 + l * a.kv_rows, b);
      }
      mbar_arrive_expect_tx(bar_v, 2 * kvc);
      for (int c = 0; c < 2; ++c)
        for (int l = 0; l < a.kv_loads; ++l)
          tma_load_3d(&tmV, bar_v, sV_u + c * kvc + l * a.kv_rows * 128, h * dh + 64 * c, key0 + l * a.kv_rows, b);
      mbar_wait(bar_qk, ph);
      tc_fence_after();
      const int n0 = NK > 256 ? 256 : NK, n1 = NK - n0;
      const uint32_t idesc0 = umma_idesc_bf16(128, n0, false, false);
      const uint32_t idesc1 = umma_idesc_bf16(128, n1 > 0 ? n1 : 16, false, false);
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t qa = sQ_u + (ks >> 2) * kChunkBytes + (ks & 3) * 32;
        const uint32_t ka = sK_u + (ks >> 2) * kvc + (ks & 3) * 32;
        umma_bf16_ss(tmem_base, umma_smem_desc_sw128(qa, 16, 1024), umma_smem_desc_sw128(ka, 16, 1024), idesc0,
                     ks > 0 ? 1u : 0u);
        if (n1 > 0)   // keys 256 .. NK-1: B rows start 256 * 128 B further, D columns start at 256
          umma_bf16_ss(tmem_base + 256u, umma_smem_desc_sw128(qa, 16, 1024),
                       umma_smem_desc_sw128(ka + 256 * 128, 16, 1024), idesc1, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, ph);
    tc_fence_after();

    const ColRange cr = my_chunks(NK, half);
    // pass 1: row max over the valid keys of this thread's column half
    float mx = -INFINITY;
    for (int c = cr.c_begin; c < cr.c_end; c += 2) {
      uint32_t v0[32], v1[32];
      const bool two = (c + 1 < cr.c_end);
      issue_chunk(trow, c * 32, NK, v0);
      if (two) issue_chunk(trow, (c + 1) * 32, NK, v1);
      tmem_ld_wait();
      mx = fmaxf(mx, chunk_max(v0, c * 32, NK, valid));
      if (two) mx = fmaxf(mx, chunk_max(v1, (c + 1) * 32, NK, valid));
    }
    red[half * 128 + r] = mx;
    __syncthreads();
    const float m_new = fmaxf(m_run, fmaxf(red[r], red[128 + r]));
    const float alpha = ex2_approx((m_run - m_new) * a.scale_log2);   // 0 on the first block (m_run = -inf)
    m_run = m_new;
    // pass 2: p = exp2((s - max) * c), row sum, bf16 P into the swizzled A-operand image (over the K block)
    sum *= alpha;
    const float mxs = m_new * a.scale_log2;
    for (int c = cr.c_begin; c < cr.c_end; ++c) {
      const int c0 = c * 32;
      uint32_t v[32];
      ld_chunk(trow, c0, NK, 0u, v);
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = ex2_approx(fmaf(__uint_as_float(v[j]), a.scale_log2, -mxs));
        pv[j] = (c0 + j < valid) ? e : 0.f;
        sum += pv[j];
      }
      const int kc = c0 >> 6, u0 = (c0 & 63) >> 3, nunits = (c0 + 32 <= NK) ? 4 : 2;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (u < nunits)
          st_shared_v4(sP_u + kc * kChunkBytes + swz_unit(r, u0 + u), pack_bf16x2(pv[8 * u + 0], pv[8 * u + 1]),
                       pack_bf16x2(pv[8 * u + 2], pv[8 * u + 3]), pack_bf16x2(pv[8 * u + 4], pv[8 * u + 5]),
                       pack_bf16x2(pv[8 * u + 6], pv[8 * u + 7]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, ph);
      // O_blk = P V : A = P (K-major over keys), B = V consumed MN-major ([keys x head_dim], head_dim
      // contiguous, the two 64-column chunks kvc bytes apart)
      const uint32_t idesc = umma_idesc_bf16(128, dh, false, true);
      const int nks = NK >> 4;
      for (int t = 0; t < nks; ++t)
        umma_bf16_ss(tmem_o, umma_smem_desc_sw128(sP_u + (t >> 2) * kChunkBytes + (t & 3) * 32, 16, 1024),
                     umma_smem_desc_sw128(sV_u + t * 2048, kvc, 1024), idesc, t > 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (multi) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int u = ub + i;
        if (u < ue) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(trow + 384u + static_cast<uint32_t>(u * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) o_acc[i][j] = fmaf(o_acc[i][j], alpha, __uint_as_float(v[j]));
        }
      }
      // the next block overwrites K / P / V and both TMEM regions: every warp has to be done reading them
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }

  red[256 + half * 128 + r] = sum;
  __syncthreads();
  sum = red[256 + r] + red[256 + 128 + r];
  const int row = row0 + r;
  {
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int u = ub + i;
      if (u < ue) {
        float f[16];
        if (multi) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = o_acc[i][j] * inv;
        } else {
          uint32_t v[16];
          tmem_ld_32x32b_x16(trow + 384u + static_cast<uint32_t>(u * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * inv;
        }
        if (row < a.N) {
          uint4* d = reinterpret_cast<uint4*>(a.o + b * a.o_bs + static_cast<long long>(row) * a.o_rs + h * dh + u * 16);
          uint4 w0, w1;
          w0.x = pack_bf16x2(f[0], f[1]);   w0.y = pack_bf16x2(f[2], f[3]);
          w0.z = pack_bf16x2(f[4], f[5]);   w0.w = pack_bf16x2(f[6], f[7]);
          w1.x = pack_bf16x2(f[8], f[9]);   w1.y = pack_bf16x2(f[10], f[11]);
          w1.z = pack_bf16x2(f[12], f[13]); w1.w = pack_bf16x2(f[14], f[15]);
          d[0] = w0;
          d[1] = w1;
        }
      }
    }
    if (row < a.N && a.lse && half == 0)
      a.lse[(static_cast<long long>(b) * a.H + h) * a.N + row] = m_run * a.scale + logf(sum);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ================================================================================================
// backward.  TMEM columns: [0,256) S then dP then dQ ; [256,384) dK (two 128-key M tiles x 64) ; [384,512) dV.
// Second generation of this kernel, restructured around what the first one's ncu source page showed
// (profiles/ncu_attn_r01.txt: ~29 % of all warp samples sat on strided global loads of D_i = rowsum(dO * O),
// 11 % on a serialised dK/dV drain, 4 % on the per-tile TMA wait; 0.277 -> 0.198 ms at the c2 shape):
//   * every Q / dO / O tile of the head (<= 2 tiles of 128 queries) is fetched by TMA at kernel start, so
//     tile 1 lands while tile 0 computes; O arrives as a third swizzled tile and D_i is reduced from
//     shared memory (two threads per row, 4 x 16-byte units each, partials exchanged through smem)
//   * dS overwrites P in place (one [128 x keys] image instead of two), which pays for the extra tiles
//   * TMEM reads are issued two chunks per wait; the dK / dV drain keeps independent register blocks
// ================================================================================================

// LONG = true: any number of tokens.  blockIdx.z picks a block of up to 256 keys; the CTA keeps that block's K, V, dK, dV
// and streams ALL query tiles past it through two Q / dO / O buffers (the tile after next is requested as soon as a tile
// retires).  dQ of a query tile is the sum over the key blocks, i.e. over CTAs: each adds its part into an fp32 buffer
// (red.global.add), which vitb_attn_bwd_tc_long converts to bf16 afterwards.  D_i is recomputed per key block from the
// O / dO tiles (cheaper than a pre-pass).  With one key block (<= 256 tokens) nothing changes: LONG = false is the
// kernel the persistent attn_bwd_ws superseded for the headline shape.
template <bool LONG>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
             const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQ,
             const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
             const __grid_constant__ AttnTc a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int kb0 = LONG ? static_cast<int>(blockIdx.z) * 256 : 0;                 // first key of this CTA's block
  const int NK = LONG ? min(256, (a.N - kb0 + 15) & ~15) : a.NK;                  // keys of the block, padded to 16
  const int kv_bytes = a.NK * 128;                                                // a.NK = rows of the K / V TMA box, always written whole
  const int mtiles = (NK + 127) >> 7;        // 128-key M tiles of dK / dV
  const int nchunks = mtiles * 2;            // the P / dS image always holds whole M tiles
  const int qtiles = (a.N + 127) >> 7;       // <= 2 unless LONG
  const int nbuf = LONG ? 2 : qtiles;        // Q / dO / O tile buffers
  uint8_t* sK = smem;
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sQ = sV + kv_bytes;               // [nbuf] tiles of 128 x 64 bf16
  uint8_t* sDO = sQ + nbuf * kChunkBytes;
  uint8_t* sO = sDO + nbuf * kChunkBytes;
  uint8_t* sP = sO + nbuf * kChunkBytes;   // P, then dS in place
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + nchunks * kChunkBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* red = reinterpret_cast<float*>(bars + 16);   // [2 halves][128 rows] partial D_i
  const uint32_t bar_kv = smem_u32(bars), bar_q = bar_kv + 8 /* [2] */, bar_s = bar_kv + 24, bar_dp = bar_kv + 32,
                 bar_dq = bar_kv + 40, bar_fin = bar_kv + 48;

  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int h = blockIdx.x, b = blockIdx.y;
  const int r = (warp & 3) * 32 + lane;
  const ColRange cr = my_chunks(NK, half);
  const uint32_t sP_u = smem_u32(sP);
  const uint32_t sQ_u = smem_u32(sQ), sDO_u = smem_u32(sDO), sO_u = smem_u32(sO), sK_u = smem_u32(sK), sV_u = smem_u32(sV);

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 7; ++i) mbar_init(bar_kv + 8 * i, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();   // first global access (the TMA loads, the LSE reads) comes after the previous grid has completed
  if (tid == 0) {
    // every operand of this head is requested now (the loads fly while TMEM is allocated and the P image
    // is zeroed); nothing is reloaded later
    mbar_arrive_expect_tx(bar_q, 3 * kChunkBytes);
    tma_load_3d(&tmQ, bar_q, sQ_u, h * DH, 0, b);
    mbar_arrive_expect_tx(bar_kv, 2 * kv_bytes);
    tma_load_3d(&tmK, bar_kv, sK_u, h * DH, kb0, b);
    tma_load_3d(&tmV, bar_kv, sV_u, h * DH, kb0, b);
    tma_load_3d(&tmDO, bar_q, sDO_u, h * DH, 0, b);
    tma_load_3d(&tmO, bar_q, sO_u, h * DH, 0, b);
    if (qtiles > 1) {
      mbar_arrive_expect_tx(bar_q + 8, 3 * kChunkBytes);
      tma_load_3d(&tmQ, bar_q + 8, sQ_u + kChunkBytes, h * DH, 128, b);
      tma_load_3d(&tmDO, bar_q + 8, sDO_u + kChunkBytes, h * DH, 128, b);
      tma_load_3d(&tmO, bar_q + 8, sO_u + kChunkBytes, h * DH, 128, b);
    }
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  // zero the P / dS image once: key columns >= NK of the last M tile are never written again
  {
    const int total16 = nchunks * kChunkBytes / 16;
    for (int i = tid; i < total16; i += kAttnThreads) st_shared_v4(sP_u + i * 16, 0u, 0u, 0u, 0u);
  }
  // LSE of this thread's row in tile 0 (tile 1's is prefetched one tile ahead)
  const float* lse_row = a.lse + (static_cast<long long>(b) * a.H + h) * a.N;
  float lse_next = (r < a.N) ? lse_row[r] : INFINITY;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();       // barrier inits, TMEM address and the zeroed image are visible to everyone
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);

  for (int qt = 0; qt < qtiles; ++qt) {
    const uint32_t ph = qt & 1;
    const int row0 = qt * 128;
    const int row = row0 + r;
    const int buf = qt % nbuf;                                  // tile buffer and the parity of its barrier
    const uint32_t bar_qt = bar_q + 8 * buf, ph_q = static_cast<uint32_t>((qt / nbuf) & 1);
    const uint32_t sQ_t = sQ_u + buf * kChunkBytes, sDO_t = sDO_u + buf * kChunkBytes, sO_t = sO_u + buf * kChunkBytes;
    const float lse2 = lse_next * 1.4426950408889634f;
    if (qt + 1 < qtiles) lse_next = (row + 128 < a.N) ? lse_row[row + 128] : INFINITY;
    if (tid == 0) {
      if (qt == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_qt, ph_q);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, NK, false, false);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)  // S = Q K^T
        umma_bf16_ss(tmem_base, umma_smem_desc_sw128(sQ_t + k * 32, 16, 1024),
                     umma_smem_desc_sw128(sK_u + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    // partial D_i = sum over this thread's 32 head-dim columns of dO * O, from the swizzled TMA tiles
    mbar_wait(bar_qt, ph_q);
    {
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t off = swz_unit(r, half * 4 + u);
        part += dot8_bf16(ld_shared_v4(sO_t + off), ld_shared_v4(sDO_t + off));
      }
      red[half * 128 + r] = part;
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    // P = exp2(S*c - LSE*log2e) -> bf16 -> sP   (rows >= N and keys >= N give exactly 0)
    auto emit_p = [&](const uint32_t (&v)[32], int c0) {
      float pv[32];   // rows >= N carry LSE = +inf, so exp2(-inf) zeroes them without a row predicate
      if (kb0 + c0 + 32 <= a.N) {
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = ex2_approx(fmaf(__uint_as_float(v[j]), a.scale_log2, -lse2));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = ex2_approx(fmaf(__uint_as_float(v[j]), a.scale_log2, -lse2));
          pv[j] = (kb0 + c0 + j < a.N) ? e : 0.f;
        }
      }
      const int kc = c0 >> 6, u0 = (c0 & 63) >> 3, nunits = (c0 + 32 <= NK) ? 4 : 2;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nunits)
          st_shared_v4(sP_u + kc * kChunkBytes + swz_unit(r, u0 + u), pack_bf16x2(pv[8 * u + 0], pv[8 * u + 1]),
                       pack_bf16x2(pv[8 * u + 2], pv[8 * u + 3]), pack_bf16x2(pv[8 * u + 4], pv[8 * u + 5]),
                       pack_bf16x2(pv[8 * u + 6], pv[8 * u + 7]));
    };
    for (int c = cr.c_begin; c < cr.c_end; c += 2) {
      uint32_t v0[32], v1[32];
      const bool two = (c + 1 < cr.c_end);
      issue_chunk(trow, c * 32, NK, v0);
      if (two) issue_chunk(trow, (c + 1) * 32, NK, v1);
      tmem_ld_wait();
      emit_p(v0, c * 32);
      if (two) emit_p(v1, (c + 1) * 32);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    const float Di = red[r] + red[128 + r];
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc_dp = umma_idesc_bf16(128, NK, false, false);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)  // dP = dO V^T  (overwrites S)
        umma_bf16_ss(tmem_base, umma_smem_desc_sw128(sDO_t + k * 32, 16, 1024),
                     umma_smem_desc_sw128(sV_u + k * 32, 16, 1024), idesc_dp, k > 0 ? 1u : 0u);
      // dV[m-tile] += P^T dO : A = P^T (MN-major image of sP), B = dO (MN-major), K = 128 query rows
      const uint32_t idesc_t = umma_idesc_bf16(128, DH, true, true);
      for (int mt = 0; mt < mtiles; ++mt)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tmem_base + 384 + mt * 64,
                       umma_smem_desc_sw128(sP_u + mt * 2 * kChunkBytes + k * 2048, kChunkBytes, 1024),
                       umma_smem_desc_sw128(sDO_t + k * 2048, 8192, 1024), idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
      umma_commit(bar_dp);   // after dV too: dS is written over P, which the dV MMAs read
    }
    __syncwarp();
    mbar_wait(bar_dp, ph);
    tc_fence_after();
    // dS / c = P * (dP - D)  -> bf16, in place over P  (the softmax scale c is applied when dQ / dK are drained)
    auto emit_ds = [&](const uint32_t (&v)[32], int c0) {
      const int kc = c0 >> 6, u0 = (c0 & 63) >> 3, nunits = (c0 + 32 <= NK) ? 4 : 2;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (u < nunits) {
          const uint32_t addr = sP_u + kc * kChunkBytes + swz_unit(r, u0 + u);
          const uint4 pp = ld_shared_v4(addr);
          const float d0 = bf16_lo(pp.x) * (__uint_as_float(v[8 * u + 0]) - Di);
          const float d1 = bf16_hi(pp.x) * (__uint_as_float(v[8 * u + 1]) - Di);
          const float d2 = bf16_lo(pp.y) * (__uint_as_float(v[8 * u + 2]) - Di);
          const float d3 = bf16_hi(pp.y) * (__uint_as_float(v[8 * u + 3]) - Di);
          const float d4 = bf16_lo(pp.z) * (__uint_as_float(v[8 * u + 4]) - Di);
          const float d5 = bf16_hi(pp.z) * (__uint_as_float(v[8 * u + 5]) - Di);
          const float d6 = bf16_lo(pp.w) * (__uint_as_float(v[8 * u + 6]) - Di);
          const float d7 = bf16_hi(pp.w) * (__uint_as_float(v[8 * u + 7]) - Di);
          st_shared_v4(addr, pack_bf16x2(d0, d1), pack_bf16x2(d2, d3), pack_bf16x2(d4, d5), pack_bf16x2(d6, d7));
        }
      }
    };
    for (int c = cr.c_begin; c < cr.c_end; c += 2) {
      uint32_t v0[32], v1[32];
      const bool two = (c + 1 < cr.c_end);
      issue_chunk(trow, c * 32, NK, v0);
      if (two) issue_chunk(trow, (c + 1) * 32, NK, v1);
      tmem_ld_wait();
      emit_ds(v0, c * 32);
      if (two) emit_ds(v1, (c + 1) * 32);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // dQ = dS K : A = dS (K-major over keys), B = K (MN-major: keys x dh)
      const uint32_t idesc_dq = umma_idesc_bf16(128, DH, false, true);
      const int nks = NK >> 4;
      for (int t = 0; t < nks; ++t)
        umma_bf16_ss(tmem_base, umma_smem_desc_sw128(sP_u + (t >> 2) * kChunkBytes + (t & 3) * 32, 16, 1024),
                     umma_smem_desc_sw128(sK_u + t * 2048, 8192, 1024), idesc_dq, t > 0 ? 1u : 0u);
      umma_commit(bar_dq);
      // dK[m-tile] += dS^T Q
      const uint32_t idesc_t = umma_idesc_bf16(128, DH, true, true);
      for (int mt = 0; mt < mtiles; ++mt)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tmem_base + 256 + mt * 64,
                       umma_smem_desc_sw128(sP_u + mt * 2 * kChunkBytes + k * 2048, kChunkBytes, 1024),
                       umma_smem_desc_sw128(sQ_t + k * 2048, 8192, 1024), idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
      umma_commit(bar_fin);
    }
    __syncwarp();
    mbar_wait(bar_dq, ph);
    tc_fence_after();
    {
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + half * 32, v);
      tmem_ld_wait();
      if constexpr (LONG) {
        // this key block's share of dQ: 32 fp32 columns of the row, added to the accumulation buffer
        if (row < a.N) {
          float* dst = a.dq_acc + (static_cast<long long>(b) * a.N + row) * (static_cast<long long>(a.H) * DH) + h * DH + half * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j]) * a.scale),
                         "f"(__uint_as_float(v[j + 1]) * a.scale), "f"(__uint_as_float(v[j + 2]) * a.scale),
                         "f"(__uint_as_float(v[j + 3]) * a.scale) : "memory");
        }
      } else {
        stage_row32_bf16(sO_t, r, half, v, a.scale);   // this tile's O buffer is dead once D_i is known
      }
    }
    fence_proxy_async_smem();
    // the next query tile overwrites sP and TMEM[0,256): wait until every MMA of this tile (dK included)
    // has retired, and until all warps have drained dQ from TMEM
    mbar_wait(bar_fin, ph);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      if constexpr (LONG) {
        if (qt + 2 < qtiles) {   // this tile's buffers are free (every MMA that read them has retired): fetch the tile after next
          mbar_arrive_expect_tx(bar_qt, 3 * kChunkBytes);
          tma_load_3d(&tmQ, bar_qt, sQ_t, h * DH, row0 + 256, b);
          tma_load_3d(&tmDO, bar_qt, sDO_t, h * DH, row0 + 256, b);
          tma_load_3d(&tmO, bar_qt, sO_t, h * DH, row0 + 256, b);
        }
      } else {   // dQ tile [128 x 64] leaves as one TMA store (rows >= N clipped)
        tma_store_3d(&tmDQ, sO_t, h * DH, row0, b);
        bulk_commit();
      }
    }
  }
  // dK, dV: TMEM lane = key within the M tile; the two threads of a lane split the 64 head-dim columns
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    if (mt < mtiles) {
      uint32_t vk[32], vv[32];
      tmem_ld_32x32b_x32(trow + 256u + static_cast<uint32_t>(mt * 64 + half * 32), vk);
      tmem_ld_32x32b_x32(trow + 384u + static_cast<uint32_t>(mt * 64 + half * 32), vv);
      tmem_ld_wait();
      // every MMA has retired, so the P / dS image is free: it stages the dK and dV tiles of both M tiles
      stage_row32_bf16(sP_u + (2 * mt) * kChunkBytes, r, half, vk, a.scale);
      stage_row32_bf16(sP_u + (2 * mt + 1) * kChunkBytes, r, half, vv, 1.0f);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    for (int mt = 0; mt < mtiles; ++mt) {
      tma_store_3d(&tmDK, sP_u + (2 * mt) * kChunkBytes, h * DH, kb0 + mt * 128, b);
      tma_store_3d(&tmDV, sP_u + (2 * mt + 1) * kChunkBytes, h * DH, kb0 + mt * 128, b);
    }
    bulk_commit();
    bulk_wait_all();   // dQ stores included: shared memory must outlive the reads
  }
  if (warp == 0) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ================================================================================================
// backward, wide heads: 64 < head_dim <= 128 (multiple of 16), any number of tokens — ViT-H/14 is head_dim 80
// (src/config.py:95-104), 257 tokens at 224 px.  Same algorithm as attn_bwd_tc<true>, re-dimensioned:
//   * the head dimension spans two 64-column 128B-swizzled chunks (the second TMA box reads 64 columns even when fewer
//     belong to the head: the K-major MMAs stop after head_dim / 16 steps, the MN-major ones take N = head_dim with the
//     chunk distance as the descriptor's leading-dimension offset — exactly as attn_fwd_tc_gen consumes V)
//   * a CTA owns a block of 128 keys (one M tile): TMEM = S / dP / dQ in [0, 128), dK in [128, 128 + head_dim),
//     dV in [256, 256 + head_dim); shared memory = K, V (2 x 2 chunks) | ONE Q / dO / O tile set (3 x 2 chunks) | P / dS
//     (2 chunks) = 192 KB, so the next query tile is requested only when the current one has retired
//   * dQ is added to the fp32 accumulation buffer, dK / dV leave as plain 16-byte global stores (no TMA-store tiles)
// ================================================================================================
struct AttnWide {
  int N, H, dh;
  float scale, scale_log2;
  float* lse;
  float* dq_acc;                       // [B, N, H*dh] fp32, zeroed by the caller
  __nv_bfloat16 *dk, *dv;
  long long dk_bs, dk_rs, dv_bs, dv_rs;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_tc_wide(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ AttnWide a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int T2 = 2 * kChunkBytes;          // one [128 x 128-column] tile: two 64-column chunks
  const int dh = a.dh;
  const int kb0 = static_cast<int>(blockIdx.z) * 128;
  const int NK = min(128, (a.N - kb0 + 15) & ~15);
  const int qtiles = (a.N + 127) >> 7;
  uint8_t* sK = smem;
  uint8_t* sV = sK + T2;
  uint8_t* sQ = sV + T2;
  uint8_t* sDO = sQ + T2;
  uint8_t* sO = sDO + T2;
  uint8_t* sP = sO + T2;                       // [128 queries x 128 keys]: P, then dS in place
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + T2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* red = reinterpret_cast<float*>(bars + 16);   // [2 halves][128 rows] partial D_i
  const uint32_t bar_kv = smem_u32(bars), bar_q = bar_kv + 8, bar_s = bar_kv + 16, bar_dp = bar_kv + 24,
                 bar_dq = bar_kv + 32, bar_fin = bar_kv + 40;

  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int h = blockIdx.x, b = blockIdx.y;
  const int r = (warp & 3) * 32 + lane;
  const ColRange cr = my_chunks(NK, half);
  const uint32_t sP_u = smem_u32(sP), sQ_u = smem_u32(sQ), sDO_u = smem_u32(sDO), sO_u = smem_u32(sO), sK_u = smem_u32(sK),
                 sV_u = smem_u32(sV);
  const int ksteps = dh >> 4;                  // 16-column steps of the head dimension
  // this thread's share of a row's head-dim columns, in 16-column units: half 0 takes the larger part
  const int u16b = half ? (ksteps + 1) / 2 : 0, u16e = half ? ksteps : (ksteps + 1) / 2;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 6; ++i) mbar_init(bar_kv + 8 * i, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  auto load_q_tile = [&](int row0) {           // Q, dO, O of one query tile: 3 tiles x 2 chunks
    mbar_arrive_expect_tx(bar_q, 3 * T2);
    for (int c = 0; c < 2; ++c) {
      tma_load_3d(&tmQ, bar_q, sQ_u + c * kChunkBytes, h * dh + 64 * c, row0, b);
      tma_load_3d(&tmDO, bar_q, sDO_u + c * kChunkBytes, h * dh + 64 * c, row0, b);
      tma_load_3d(&tmO, bar_q, sO_u + c * kChunkBytes, h * dh + 64 * c, row0, b);
    }
  };
  if (tid == 0) {
    load_q_tile(0);
    mbar_arrive_expect_tx(bar_kv, 2 * T2);
    for (int c = 0; c < 2; ++c) {
      tma_load_3d(&tmK, bar_kv, sK_u + c * kChunkBytes, h * dh + 64 * c, kb0, b);
      tma_load_3d(&tmV, bar_kv, sV_u + c * kChunkBytes, h * dh + 64 * c, kb0, b);
    }
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  for (int i = tid; i < T2 / 16; i += kAttnThreads) st_shared_v4(sP_u + i * 16, 0u, 0u, 0u, 0u);   // key columns >= NK stay zero
  const float* lse_row = a.lse + (static_cast<long long>(b) * a.H + h) * a.N;
  float lse_next = (r < a.N) ? lse_row[r] : INFINITY;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t T_DK = 128u, T_DV = 256u;

  for (int qt = 0; qt < qtiles; ++qt) {
    const uint32_t ph = qt & 1;
    const int row0 = qt * 128;
    const int row = row0 + r;
    const float lse2 = lse_next * 1.4426950408889634f;
    if (qt + 1 < qtiles) lse_next = (row + 128 < a.N) ? lse_row[row + 128] : INFINITY;
    if (tid == 0) {
      if (qt == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, ph);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, NK, false, false);
      for (int ks = 0; ks < ksteps; ++ks)      // S = Q K^T over head_dim / 16 steps (chunk ks / 4)
        umma_bf16_ss(tmem_base, umma_smem_desc_sw128(sQ_u + (ks >> 2) * kChunkBytes + (ks & 3) * 32, 16, 1024),
                     umma_smem_desc_sw128(sK_u + (ks >> 2) * kChunkBytes + (ks & 3) * 32, 16, 1024), idesc, ks > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_q, ph);
    {   // partial D_i over this thread's share of the head-dim columns
      float part = 0.f;
      for (int u = 2 * u16b; u < 2 * u16e; ++u) {
        const uint32_t off = static_cast<uint32_t>((u >> 3) * kChunkBytes) + swz_unit(r, u & 7);
        part += dot8_bf16(ld_shared_v4(sO_u + off), ld_shared_v4(sDO_u + off));
      }
      red[half * 128 + r] = part;
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    auto emit_p = [&](const uint32_t (&v)[32], int c0) {
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = ex2_approx(fmaf(__uint_as_float(v[j]), a.scale_log2, -lse2));
        pv[j] = (kb0 + c0 + j < a.N) ? e : 0.f;
      }
      const int kc = c0 >> 6, u0 = (c0 & 63) >> 3, nunits = (c0 + 32 <= NK) ? 4 : 2;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nunits)
          st_shared_v4(sP_u + kc * kChunkBytes + swz_unit(r, u0 + u), pack_bf16x2(pv[8 * u + 0], pv[8 * u + 1]),
                       pack_bf16x2(pv[8 * u + 2], pv[8 * u + 3]), pack_bf16x2(pv[8 * u + 4], pv[8 * u + 5]),
                       pack_bf16x2(pv[8 * u + 6], pv[8 * u + 7]));
    };
    for (int c = cr.c_begin; c < cr.c_end; ++c) {
      uint32_t v0[32];
      issue_chunk(trow, c * 32, NK, v0);
      tmem_ld_wait();
      emit_p(v0, c * 32);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    const float Di = red[r] + red[128 + r];
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc_dp = umma_idesc_bf16(128, NK, false, false);
      for (int ks = 0; ks < ksteps; ++ks)      // dP = dO V^T (overwrites S)
        umma_bf16_ss(tmem_base, umma_smem_desc_sw128(sDO_u + (ks >> 2) * kChunkBytes + (ks & 3) * 32, 16, 1024),
                     umma_smem_desc_sw128(sV_u + (ks >> 2) * kChunkBytes + (ks & 3) * 32, 16, 1024), idesc_dp, ks > 0 ? 1u : 0u);
      // dV += P^T dO : A = P^T (MN-major image of sP: 128 keys = two 64-key chunks), B = dO (MN-major, N = head_dim
      // across the two column chunks), K = 128 query rows in 8 steps
      const uint32_t idesc_t = umma_idesc_bf16(128, dh, true, true);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16_ss(tmem_base + T_DV, umma_smem_desc_sw128(sP_u + k * 2048, kChunkBytes, 1024),
                     umma_smem_desc_sw128(sDO_u + k * 2048, kChunkBytes, 1024), idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
      umma_commit(bar_dp);
    }
    __syncwarp();
    mbar_wait(bar_dp, ph);
    tc_fence_after();
    auto emit_ds = [&](const uint32_t (&v)[32], int c0) {
      const int kc = c0 >> 6, u0 = (c0 & 63) >> 3, nunits = (c0 + 32 <= NK) ? 4 : 2;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (u < nunits) {
          const uint32_t addr = sP_u + kc * kChunkBytes + swz_unit(r, u0 + u);
          const uint4 pp = ld_shared_v4(addr);
          const float d0 = bf16_lo(pp.x) * (__uint_as_float(v[8 * u + 0]) - Di);
          const float d1 = bf16_hi(pp.x) * (__uint_as_float(v[8 * u + 1]) - Di);
          const float d2 = bf16_lo(pp.y) * (__uint_as_float(v[8 * u + 2]) - Di);
          const float d3 = bf16_hi(pp.y) * (__uint_as_float(v[8 * u + 3]) - Di);
          const float d4 = bf16_lo(pp.z) * (__uint_as_float(v[8 * u + 4]) - Di);
          const float d5 = bf16_hi(pp.z) * (__uint_as_float(v[8 * u + 5]) - Di);
          const float d6 = bf16_lo(pp.w) * (__uint_as_float(v[8 * u + 6]) - Di);
          const float d7 = bf16_hi(pp.w) * (__uint_as_float(v[8 * u + 7]) - Di);
          st_shared_v4(addr, pack_bf16x2(d0, d1), pack_bf16x2(d2, d3), pack_bf16x2(d4, d5), pack_bf16x2(d6, d7));
        }
      }
    };
    for (int c = cr.c_begin; c < cr.c_end; ++c) {
      uint32_t v0[32];
      issue_chunk(trow, c * 32, NK, v0);
      tmem_ld_wait();
      emit_ds(v0, c * 32);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // dQ = dS K : A = dS (K-major over keys), B = K (MN-major: keys x head_dim, chunk distance kChunkBytes)
      const uint32_t idesc_dq = umma_idesc_bf16(128, dh, false, true);
      const int nks = NK >> 4;
      for (int t = 0; t < nks; ++t)
        umma_bf16_ss(tmem_base, umma_smem_desc_sw128(sP_u + (t >> 2) * kChunkBytes + (t & 3) * 32, 16, 1024),
                     umma_smem_desc_sw128(sK_u + t * 2048, kChunkBytes, 1024), idesc_dq, t > 0 ? 1u : 0u);
      umma_commit(bar_dq);
      const uint32_t idesc_t = umma_idesc_bf16(128, dh, true, true);
#pragma unroll
      for (int k = 0; k < 8; ++k)              // dK += dS^T Q
        umma_bf16_ss(tmem_base + T_DK, umma_smem_desc_sw128(sP_u + k * 2048, kChunkBytes, 1024),
                     umma_smem_desc_sw128(sQ_u + k * 2048, kChunkBytes, 1024), idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
      umma_commit(bar_fin);
    }
    __syncwarp();
    mbar_wait(bar_dq, ph);
    tc_fence_after();
    for (int u = u16b; u < u16e; ++u) {        // this key block's share of dQ, 16 columns at a time
      uint32_t v[16];
      tmem_ld_32x32b_x16(trow + static_cast<uint32_t>(u * 16), v);
      tmem_ld_wait();
      if (row < a.N) {
        float* dst = a.dq_acc + (static_cast<long long>(b) * a.N + row) * (static_cast<long long>(a.H) * dh) + h * dh + u * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j]) * a.scale),
                       "f"(__uint_as_float(v[j + 1]) * a.scale), "f"(__uint_as_float(v[j + 2]) * a.scale),
                       "f"(__uint_as_float(v[j + 3]) * a.scale) : "memory");
      }
    }
    mbar_wait(bar_fin, ph);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0 && qt + 1 < qtiles) load_q_tile(row0 + 128);   // the tile set is free: every MMA that read it has retired
  }
  // dK, dV: TMEM lane = key of the block; 16-column units split between the two threads of a lane
  {
    const int key = kb0 + r;
    for (int u = u16b; u < u16e; ++u) {
      uint32_t vk[16], vv[16];
      tmem_ld_32x32b_x16(trow + T_DK + static_cast<uint32_t>(u * 16), vk);
      tmem_ld_32x32b_x16(trow + T_DV + static_cast<uint32_t>(u * 16), vv);
      tmem_ld_wait();
      if (key < a.N) {
        uint4* dk = reinterpret_cast<uint4*>(a.dk + b * a.dk_bs + static_cast<long long>(key) * a.dk_rs + h * dh + u * 16);
        uint4* dv = reinterpret_cast<uint4*>(a.dv + b * a.dv_bs + static_cast<long long>(key) * a.dv_rs + h * dh + u * 16);
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          uint4 ok, ov;
          ok.x = pack_bf16x2(__uint_as_float(vk[8 * q4 + 0]) * a.scale, __uint_as_float(vk[8 * q4 + 1]) * a.scale);
          ok.y = pack_bf16x2(__uint_as_float(vk[8 * q4 + 2]) * a.scale, __uint_as_float(vk[8 * q4 + 3]) * a.scale);
          ok.z = pack_bf16x2(__uint_as_float(vk[8 * q4 + 4]) * a.scale, __uint_as_float(vk[8 * q4 + 5]) * a.scale);
          ok.w = pack_bf16x2(__uint_as_float(vk[8 * q4 + 6]) * a.scale, __uint_as_float(vk[8 * q4 + 7]) * a.scale);
          ov.x = pack_bf16x2(__uint_as_float(vv[8 * q4 + 0]), __uint_as_float(vv[8 * q4 + 1]));
          ov.y = pack_bf16x2(__uint_as_float(vv[8 * q4 + 2]), __uint_as_float(vv[8 * q4 + 3]));
          ov.z = pack_bf16x2(__uint_as_float(vv[8 * q4 + 4]), __uint_as_float(vv[8 * q4 + 5]));
          ov.w = pack_bf16x2(__uint_as_float(vv[8 * q4 + 6]), __uint_as_float(vv[8 * q4 + 7]));
          dk[q4] = ok;
          dv[q4] = ov;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int make_head_map(CUtensorMap* m, const void* base, int H, int N, int B, long long row_stride, long long batch_stride,
                  int box_rows, int dh = DH) {
  uint64_t dims[3] = {(uint64_t)H * dh, (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
  uint32_t box[3] = {DH, (uint32_t)box_rows, 1};
  return vitb_make_tmap_nd_bf16(m, base, 3, dims, str, box);
}

bool gen_ok(int head_dim, int Nq, int Nk) {   // attn_fwd_tc_gen: any head_dim in [64, 128] % 16, any token count
  return head_dim >= DH && head_dim <= 128 && head_dim % 16 == 0 && Nq == Nk && Nk >= 1;
}

int check_common(const vitb_attn_params* p, const char* who, bool allow_wide = false) {
  VITB_REQUIRE(p && p->struct_bytes == (int)sizeof(vitb_attn_params), VITB_ERR_BAD_ARG, "%s: ABI mismatch", who);
  VITB_REQUIRE(p->dtype == VITB_BF16, VITB_ERR_UNSUPPORTED_SHAPE, "%s: bf16 only", who);
  if (allow_wide && gen_ok(p->head_dim, p->Nq, p->Nk)) {
    VITB_REQUIRE(p->q && p->k && p->v && p->o, VITB_ERR_BAD_ARG, "%s: null tensor", who);
    VITB_REQUIRE(p->o_row_stride % 8 == 0 && p->o_batch_stride % 8 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "%s: o strides %% 8", who);
    return VITB_OK;
  }
  VITB_REQUIRE(p->head_dim == DH, VITB_ERR_UNSUPPORTED_SHAPE, "%s: head_dim %d (only 64)", who, p->head_dim);
  VITB_REQUIRE(p->Nq == p->Nk && p->Nk >= 1 && p->Nk <= 256, VITB_ERR_UNSUPPORTED_SHAPE,
               "%s: Nq=%d Nk=%d (need Nq == Nk <= 256)", who, p->Nq, p->Nk);
  VITB_REQUIRE(p->q && p->k && p->v && p->o, VITB_ERR_BAD_ARG, "%s: null tensor", who);
  VITB_REQUIRE(p->o_row_stride % 8 == 0 && p->o_batch_stride % 8 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "%s: o strides %% 8", who);
  return VITB_OK;
}

}  // namespace

extern "C" int vitb_attn_supported_tc(int head_dim, int Nq, int Nk) {
  return head_dim == DH && Nq == Nk && Nk >= 1 && Nk <= 256;
}

extern "C" int vitb_attn_fwd_supported_tc(int head_dim, int Nq, int Nk) {
  return vitb_attn_supported_tc(head_dim, Nq, Nk) || gen_ok(head_dim, Nq, Nk);
}

namespace {
int launch_fwd_gen(const vitb_attn_params* p, cudaStream_t stream) {
  const int N = p->Nk, dh = p->head_dim;
  const int NKall = (N + 15) & ~15;
  AttnGen a{};
  a.N = N; a.H = p->H; a.dh = dh;
  a.KB = NKall < 320 ? NKall : 320;
  a.nblocks = (N + a.KB - 1) / a.KB;
  a.kv_loads = a.KB > 256 ? 2 : 1;
  a.kv_rows = a.KB / a.kv_loads;        // KB is a multiple of 16, so halves stay multiples of 8 (swizzle atom)
  a.scale = 1.0f / sqrtf((float)dh);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.o = reinterpret_cast<__nv_bfloat16*>(p->o); a.o_bs = p->o_batch_stride; a.o_rs = p->o_row_stride;
  a.lse = p->lse;
  int st;
  CUtensorMap tq, tk, tv;
  if ((st = make_head_map(&tq, p->q, p->H, N, p->B, p->q_row_stride, p->q_batch_stride, 128, dh)) != VITB_OK) return st;
  if ((st = make_head_map(&tk, p->k, p->H, N, p->B, p->k_row_stride, p->k_batch_stride, a.kv_rows, dh)) != VITB_OK) return st;
  if ((st = make_head_map(&tv, p->v, p->H, N, p->B, p->v_row_stride, p->v_batch_stride, a.kv_rows, dh)) != VITB_OK) return st;
  const int kvc = a.KB * 128, npchunks = (a.KB + 63) / 64;
  const int kp_bytes = 2 * kvc > npchunks * kChunkBytes ? 2 * kvc : npchunks * kChunkBytes;
  // V block | Q | K block / P | barriers + TMEM slot | row statistics | alignment slack
  const int smem = 2 * kvc + 2 * kChunkBytes + kp_bytes + 64 + 4 * 128 * 4 + 1024;
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_fwd_tc_gen: %d B of shared memory", smem);
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((N + 127) / 128, p->H, p->B);
  VITB_CUDA_CHECK(vitb_launch<kPdlAttn>(attn_fwd_tc_gen, grid, dim3(kAttnThreads), smem, stream, tq, tk, tv, a));
  VITB_LAUNCH_CHECK("attn_fwd_tc_gen");
  return VITB_OK;
}
}  // namespace

extern "C" int vitb_attn_fwd_tc(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  st = check_common(p, "attn_fwd_tc", true);
  if (st != VITB_OK) return st;
  if (p->B == 0) return VITB_OK;
  if (p->head_dim != DH || p->Nk > 256) return launch_fwd_gen(p, reinterpret_cast<cudaStream_t>(stream_));
  const int N = p->Nk, NK = (N + 15) & ~15;
  CUtensorMap tq, tk, tv;
  if ((st = make_head_map(&tq, p->q, p->H, N, p->B, p->q_row_stride, p->q_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tk, p->k, p->H, N, p->B, p->k_row_stride, p->k_batch_stride, NK)) != VITB_OK) return st;
  if ((st = make_head_map(&tv, p->v, p->H, N, p->B, p->v_row_stride, p->v_batch_stride, NK)) != VITB_OK) return st;
  AttnTc a{};
  a.N = N; a.NK = NK; a.H = p->H;
  a.scale = 1.0f / sqrtf((float)DH);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.lse = p->lse;
  const int kv_bytes = NK * 128, nchunks = (NK + 63) / 64;
  const int u_bytes = (kChunkBytes + kv_bytes) > nchunks * kChunkBytes ? (kChunkBytes + kv_bytes) : nchunks * kChunkBytes;
  const int smem = kv_bytes + u_bytes + 64 + 4 * 128 * 4 + 1024;
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((N + 127) / 128, p->H, p->B);
  CUtensorMap to;
  if ((st = make_head_map(&to, p->o, p->H, N, p->B, p->o_row_stride, p->o_batch_stride, 128)) != VITB_OK) return st;
  VITB_CUDA_CHECK(vitb_launch<kPdlAttn>(attn_fwd_tc, grid, dim3(kAttnThreads), smem, reinterpret_cast<cudaStream_t>(stream_), tq, tk,
                              tv, to, a));
  VITB_LAUNCH_CHECK("attn_fwd_tc");
  return VITB_OK;
}

extern "C" int vitb_attn_bwd_tc(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  st = check_common(p, "attn_bwd_tc");
  if (st != VITB_OK) return st;
  if (p->B == 0) return VITB_OK;
  VITB_REQUIRE(p->lse && p->dout && p->dq && p->dk && p->dv, VITB_ERR_BAD_ARG, "attn_bwd_tc: null tensor");
  VITB_REQUIRE(p->dq_row_stride % 8 == 0 && p->dk_row_stride % 8 == 0 && p->dv_row_stride % 8 == 0 &&
                   p->do_row_stride % 8 == 0,
               VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_tc: gradient row strides %% 8");
  const int N = p->Nk, NK = (N + 15) & ~15;
  CUtensorMap tq, tk, tv, tdo;
  if ((st = make_head_map(&tq, p->q, p->H, N, p->B, p->q_row_stride, p->q_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tk, p->k, p->H, N, p->B, p->k_row_stride, p->k_batch_stride, NK)) != VITB_OK) return st;
  if ((st = make_head_map(&tv, p->v, p->H, N, p->B, p->v_row_stride, p->v_batch_stride, NK)) != VITB_OK) return st;
  if ((st = make_head_map(&tdo, p->dout, p->H, N, p->B, p->do_row_stride, p->do_batch_stride, 128)) != VITB_OK) return st;
  AttnTc a{};
  a.N = N; a.NK = NK; a.H = p->H;
  a.scale = 1.0f / sqrtf((float)DH);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.lse = p->lse;
  CUtensorMap to;
  if ((st = make_head_map(&to, p->o, p->H, N, p->B, p->o_row_stride, p->o_batch_stride, 128)) != VITB_OK) return st;
  const int kv_bytes = NK * 128, mtiles = (NK + 127) / 128, nchunks = 2 * mtiles, qtiles = (N + 127) / 128;
  // K, V | Q, dO, O tiles of every query tile | one P/dS image | barriers + TMEM slot | D_i partials | alignment slack
  const int smem = 2 * kv_bytes + 3 * qtiles * kChunkBytes + nchunks * kChunkBytes + 128 + 1024 + 1024;
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_tc: %d B of shared memory", smem);
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(p->H, p->B);
  VITB_REQUIRE(p->dq_batch_stride % 8 == 0 && p->dk_batch_stride % 8 == 0 && p->dv_batch_stride % 8 == 0,
               VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_tc: gradient batch strides %% 8");
  CUtensorMap tdq, tdk, tdv;
  if ((st = make_head_map(&tdq, p->dq, p->H, N, p->B, p->dq_row_stride, p->dq_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tdk, p->dk, p->H, N, p->B, p->dk_row_stride, p->dk_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tdv, p->dv, p->H, N, p->B, p->dv_row_stride, p->dv_batch_stride, 128)) != VITB_OK) return st;
  VITB_CUDA_CHECK(vitb_launch<kPdlAttn>(attn_bwd_tc<false>, grid, dim3(kAttnThreads), smem, reinterpret_cast<cudaStream_t>(stream_), tq, tk,
                              tv, tdo, to, tdq, tdk, tdv, a));
  VITB_LAUNCH_CHECK("attn_bwd_tc");
  return VITB_OK;
}

// fp32 dQ accumulation buffer [rows, cols] -> bf16 dq through its strides (rows = B*N, cols = H*64)
namespace {
__global__ void __launch_bounds__(256)
dq_to_bf16_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int N, int cols, long long dq_bs, long long dq_rs,
                  long long total4) {
  const int c4 = cols >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / c4;
    const int c = static_cast<int>(i - row * c4) * 4;
    const float4 v = reinterpret_cast<const float4*>(acc)[i];
    const long long bimg = row / N, n = row - bimg * N;
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dq + bimg * dq_bs + n * dq_rs + c) = o;
  }
}
}  // namespace

extern "C" int vitb_attn_bwd_long_supported(int head_dim, int Nq, int Nk) {
  return head_dim >= DH && head_dim <= 128 && head_dim % 16 == 0 && Nq == Nk && Nk >= 1;
}

namespace {
int launch_bwd_wide(const vitb_attn_params* p, float* dq_acc, cudaStream_t stream) {
  const int N = p->Nk, dh = p->head_dim;
  int st;
  CUtensorMap tq, tk, tv, tdo, to;
  if ((st = make_head_map(&tq, p->q, p->H, N, p->B, p->q_row_stride, p->q_batch_stride, 128, dh)) != VITB_OK) return st;
  if ((st = make_head_map(&tk, p->k, p->H, N, p->B, p->k_row_stride, p->k_batch_stride, 128, dh)) != VITB_OK) return st;
  if ((st = make_head_map(&tv, p->v, p->H, N, p->B, p->v_row_stride, p->v_batch_stride, 128, dh)) != VITB_OK) return st;
  if ((st = make_head_map(&tdo, p->dout, p->H, N, p->B, p->do_row_stride, p->do_batch_stride, 128, dh)) != VITB_OK) return st;
  if ((st = make_head_map(&to, p->o, p->H, N, p->B, p->o_row_stride, p->o_batch_stride, 128, dh)) != VITB_OK) return st;
  AttnWide a{};
  a.N = N; a.H = p->H; a.dh = dh;
  a.scale = 1.0f / sqrtf((float)dh);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.lse = p->lse;
  a.dq_acc = dq_acc;
  a.dk = reinterpret_cast<__nv_bfloat16*>(p->dk); a.dk_bs = p->dk_batch_stride; a.dk_rs = p->dk_row_stride;
  a.dv = reinterpret_cast<__nv_bfloat16*>(p->dv); a.dv_bs = p->dv_batch_stride; a.dv_rs = p->dv_row_stride;
  // K, V | Q, dO, O | P / dS: six [128 x 128-column] tiles | barriers + TMEM slot | D_i partials | alignment slack
  const int smem = 6 * 2 * kChunkBytes + 128 + 1024 + 1024;
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(p->H, p->B, (N + 127) / 128);
  VITB_CUDA_CHECK(vitb_launch<kPdlAttn>(attn_bwd_tc_wide, grid, dim3(kAttnThreads), smem, stream, tq, tk, tv, tdo, to, a));
  VITB_LAUNCH_CHECK("attn_bwd_tc_wide");
  return VITB_OK;
}
}  // namespace

// Backward for ANY number of tokens (head_dim 64): 384 px fine-tuning gives 577 tokens (src/config.py:12).  dq_acc is an
// fp32 [B, N, H*64] scratch buffer that the CALLER has zeroed; the key-block CTAs add their shares of dQ into it and a
// second kernel writes p->dq (bf16) from it.  dk / dv are written directly.
extern "C" int vitb_attn_bwd_tc_long(const vitb_attn_params* p, float* dq_acc, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(p && p->struct_bytes == (int)sizeof(vitb_attn_params), VITB_ERR_BAD_ARG, "attn_bwd_tc_long: ABI mismatch");
  VITB_REQUIRE(p->dtype == VITB_BF16 && vitb_attn_bwd_long_supported(p->head_dim, p->Nq, p->Nk), VITB_ERR_UNSUPPORTED_SHAPE,
               "attn_bwd_tc_long: bf16, head_dim 64..128 (multiple of 16), Nq == Nk (dh=%d Nq=%d Nk=%d)", p->head_dim, p->Nq, p->Nk);
  if (p->B == 0) return VITB_OK;
  VITB_REQUIRE(p->q && p->k && p->v && p->o && p->lse && p->dout && p->dq && p->dk && p->dv && dq_acc, VITB_ERR_BAD_ARG,
               "attn_bwd_tc_long: null tensor");
  VITB_REQUIRE(((reinterpret_cast<uintptr_t>(dq_acc) | reinterpret_cast<uintptr_t>(p->dq) | reinterpret_cast<uintptr_t>(p->dk) |
                 reinterpret_cast<uintptr_t>(p->dv)) & 15u) == 0,
               VITB_ERR_BAD_ARG, "attn_bwd_tc_long: dq_acc, dq, dk, dv must be 16-byte aligned");
  const long long strides[] = {p->q_row_stride, p->k_row_stride, p->v_row_stride, p->o_row_stride, p->do_row_stride,
                               p->dq_row_stride, p->dk_row_stride, p->dv_row_stride, p->q_batch_stride, p->k_batch_stride,
                               p->v_batch_stride, p->o_batch_stride, p->do_batch_stride, p->dq_batch_stride, p->dk_batch_stride,
                               p->dv_batch_stride};
  for (long long sv : strides) VITB_REQUIRE(sv % 8 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_tc_long: strides must be multiples of 8");
  const int N = p->Nk;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int cols = p->H * p->head_dim;
  const long long total4 = static_cast<long long>(p->B) * N * (cols / 4);
  long long blocks = (total4 + 255) / 256;
  const long long cap = static_cast<long long>(vitb_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (p->head_dim != DH) {       // wide heads: 128-key blocks, two head-dim chunks
    if ((st = launch_bwd_wide(p, dq_acc, stream)) != VITB_OK) return st;
    dq_to_bf16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(p->dq), N, cols,
                                                                    p->dq_batch_stride, p->dq_row_stride, total4);
    VITB_LAUNCH_CHECK("dq_to_bf16_kernel");
    return VITB_OK;
  }
  const int kv_box = N >= 256 ? 256 : ((N + 15) & ~15);
  CUtensorMap tq, tk, tv, tdo, to, tdq, tdk, tdv;
  if ((st = make_head_map(&tq, p->q, p->H, N, p->B, p->q_row_stride, p->q_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tk, p->k, p->H, N, p->B, p->k_row_stride, p->k_batch_stride, kv_box)) != VITB_OK) return st;
  if ((st = make_head_map(&tv, p->v, p->H, N, p->B, p->v_row_stride, p->v_batch_stride, kv_box)) != VITB_OK) return st;
  if ((st = make_head_map(&tdo, p->dout, p->H, N, p->B, p->do_row_stride, p->do_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&to, p->o, p->H, N, p->B, p->o_row_stride, p->o_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tdq, p->dq, p->H, N, p->B, p->dq_row_stride, p->dq_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tdk, p->dk, p->H, N, p->B, p->dk_row_stride, p->dk_batch_stride, 128)) != VITB_OK) return st;
  if ((st = make_head_map(&tdv, p->dv, p->H, N, p->B, p->dv_row_stride, p->dv_batch_stride, 128)) != VITB_OK) return st;
  AttnTc a{};
  a.N = N; a.NK = kv_box; a.H = p->H;
  a.scale = 1.0f / sqrtf((float)DH);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.lse = p->lse;
  a.dq_acc = dq_acc;
  // K, V boxes (256 keys each) | two Q / dO / O tile buffers | the P / dS image (4 chunks) | barriers + TMEM slot | D_i | slack
  const int smem = 2 * kv_box * 128 + 3 * 2 * kChunkBytes + 4 * kChunkBytes + 128 + 1024 + 1024;
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_tc_long: %d B of shared memory", smem);
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(p->H, p->B, (N + 255) / 256);
  VITB_CUDA_CHECK(vitb_launch<kPdlAttn>(attn_bwd_tc<true>, grid, dim3(kAttnThreads), smem, stream, tq, tk, tv, tdo, to, tdq, tdk, tdv, a));
  VITB_LAUNCH_CHECK("attn_bwd_tc<long>");
  dq_to_bf16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(p->dq), N, cols,
                                                                  p->dq_batch_stride, p->dq_row_stride, total4);
  VITB_LAUNCH_CHECK("dq_to_bf16_kernel");
  return VITB_OK;
}
