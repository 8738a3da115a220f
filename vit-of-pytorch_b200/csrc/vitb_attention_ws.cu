// vitb_attention_ws.cu — persistent, warp-specialised SelfAttention forward / backward on tcgen05 (bf16 in, fp32
// in TMEM) for head_dim 64 and up to 256 tokens (ViT-B/L at 224 px: N = 197 or 50).
//
// Replaces the q k^T / softmax / v chain of SelfAttention.forward (src/model.py:90-97) and Attention.forward
// (res-vit/model.py:273-293) plus its autograd backward — same contract as vitb_attn_fwd_tc / vitb_attn_bwd_tc.
//
// Why a second generation: the one-CTA-per-tile kernels of vitb_attention_tc.cu run every phase of their chain
// (TMA -> MMA -> CUDA-core pass -> CTA barrier -> MMA -> store) exposed; ncu showed 12-14 % tensor-pipe activity and
// 23-46 % issue utilisation (profiles/ncu_r01c.txt), i.e. latency-bound.  Here ONE CTA per SM stays resident and
// walks a contiguous range of work items; a TMA producer warp, a single-thread tcgen05 issuer warp and the
// CUDA-core warps run as a pipeline connected by mbarriers, so loads, MMAs and the softmax arithmetic of
// neighbouring items overlap:
//
//   forward  (attn_fwd_ws)  item = (image, head, 128-query tile).  Two softmax groups of 8 warps (two threads per
//            query row, each half of the keys) alternate over the items; each group owns a 256-column TMEM buffer
//            (S, later O in its first 64 columns) and a P image in shared memory.  K / V are loaded once per head and
//            shared by its tiles.  While group g does the softmax of item i, the issuer has S(i+1) in flight for the
//            other group and P(i-1) V behind it.
//   backward (attn_bwd_ws)  item = (image, head); iterations (key tile kt, query tile qt), kt outer.  TMEM: S | dP |
//            dQ[2 query tiles] | dK | dV = 512 columns.  Eight warps (two threads per query row) turn S into P and
//            dP into dS; the issuer runs S(i+1) and dV(i) behind P(i), dP(i+1), dK(i), dQ(i) behind dS(i); finished
//            accumulators (dV, dK after a key tile; dQ after a head) are drained one pipeline slot later, through one
//            staging tile and TMA stores, so nobody waits for an MMA group that was only just issued.
#include "../../include/vitb200.h"
#include <stdlib.h>

#include "vitb_common.cuh"
#include "vitb_attn_util.cuh"

namespace {
using namespace vitb;
using namespace vitb::attn;

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

struct WsParams {
  int N;          // tokens (queries == keys)
  int NK;         // keys padded to a multiple of 16 (forward: extent of S / P)
  int H;
  int qtiles;     // 128-query tiles per head (1 or 2)
  int ktiles;     // 128-key tiles per head (backward)
  int total;      // work items: forward (head, query tile) pairs; backward heads
  int s_stride;   // forward: TMEM columns between the two groups' S buffers
  int o_col;      // forward: TMEM column of the shared O accumulator, or -1: O overwrites the first 64 columns of S
  float scale;       // 1/sqrt(dh)
  float scale_log2;  // scale * log2(e)
  float* lse;     // [B,H,N]
};

// UMMA shared-memory descriptors are built once per buffer; k-steps advance the 14-bit start-address field (16-byte
// units; shared memory ends below 256 KiB, so the field cannot carry into its neighbours)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return umma_smem_desc_sw128(addr, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo) { return umma_smem_desc_sw128(addr, lbo, 1024); }

// ================================================================================================
// forward
// ================================================================================================
constexpr int kFwdThreads = 576;         // warps 0-15: two softmax groups of 8; warp 16: TMA producer; warp 17: tcgen05 issuer
enum FwdBar { FB_K = 0, FB_V = 1, FB_Q = 2 /*[2]*/, FB_S = 4 /*[2]*/, FB_P = 6 /*[2] 256*/, FB_O = 8 /*[2]*/, FB_D = 10 /*[2] 256*/, FB_COUNT = 12 };

// Padding keys (columns N .. NK-1, at most 15, all inside the last 16-key unit) are switched off by ADDING a register
// vector of 0 / -inf to the scores of that unit: exp2(-inf) = 0 and max(-inf, .) ignore them without a compare + select
// per element (the first version of this kernel spent as many instructions on its one masked chunk as on seven plain ones).
__device__ __forceinline__ void add_mask16(uint32_t* v, const uint64_t (&am)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x0, x1;
    upk2(add2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), am[j]), x0, x1);
    v[2 * j] = __float_as_uint(x0);
    v[2 * j + 1] = __float_as_uint(x1);
  }
}

// W (16 or 32) score columns of this thread's row, already in registers (padding already at -inf) ->
// p = exp2(s*c - m*c) -> row-sum partial, bf16 -> the 128B-swizzled P image (keys c .. c+W-1 of row r).
template <int W>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&v)[32], uint64_t sc2, uint64_t nm2, uint64_t& sum2a,
                                              uint64_t& sum2b, uint32_t image, int r, int rx, int c) {
  uint32_t pk[W / 2];
#pragma unroll
  for (int j = 0; j < W / 2; ++j) {
    float x0, x1;
    upk2(fma2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc2, nm2), x0, x1);
    const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
    const uint64_t e2 = pk2(e0, e1);
    if (j & 1) sum2b = add2(sum2b, e2); else sum2a = add2(sum2a, e2);
    pk[j] = pack_bf16x2(e0, e1);
  }
  // 16 keys (two 16-byte units) never straddle a 64-key chunk of the image, 32 keys may (c is a multiple of 16)
#pragma unroll
  for (int q = 0; q < W / 16; ++q) {
    const int cq = c + 16 * q;
    const uint32_t row_addr = image + static_cast<uint32_t>((cq >> 6) * kChunkBytes + r * 128);
    const int u8 = (cq & 63) >> 3;
    st_shared_v4(row_addr + static_cast<uint32_t>((u8 ^ rx) << 4), pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3]);
    st_shared_v4(row_addr + static_cast<uint32_t>(((u8 + 1) ^ rx) << 4), pk[8 * q + 4], pk[8 * q + 5], pk[8 * q + 6], pk[8 * q + 7]);
  }
}

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_ws(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
            const __grid_constant__ WsParams a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need 1024-byte aligned bases
  const int NK = a.NK;
  const int kv_bytes = NK * 128;                 // [NK keys x 64] bf16, 128B-swizzled rows
  const int pchunks = (NK + 63) >> 6;            // 64-key chunks of a P image
  const int p_bytes = pchunks * kChunkBytes;
  const uint32_t sK_u = smem_u32(smem);
  const uint32_t sV_u = sK_u + kv_bytes;
  const uint32_t sQ_u = sV_u + kv_bytes;         // [2] query tiles, one per softmax group
  const uint32_t sP_u = sQ_u + 2 * kChunkBytes;  // [2] P images; chunk 0 doubles as the O staging tile
  uint8_t* tail = smem + 2 * kv_bytes + 2 * kChunkBytes + 2 * p_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + FB_COUNT);
  float* red = reinterpret_cast<float*>(bars + FB_COUNT + 2);   // [2 groups][max | sum][2 halves][128 rows]
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * static_cast<uint32_t>(i); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int i = 0; i < FB_COUNT; ++i) mbar_init(bar(i), ((i >= FB_P && i < FB_O) || i >= FB_D) ? 256 : 1);
    fence_barrier_init();
  }
  if (warp == 17) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // first global access comes after the previous grid has completed

  // this CTA's contiguous range of (head, query tile) items
  const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * a.total / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * a.total / gridDim.x);
  const int n = t_end - t_begin;
  const int QT = a.qtiles;

  if (warp == 16) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int q_issued = 0;
      auto wait_s = [&](int i) { mbar_wait_backoff(bar(FB_S + (i & 1)), static_cast<uint32_t>((i >> 1) & 1), 64); };
      auto wait_o = [&](int i) { mbar_wait_backoff(bar(FB_O + (i & 1)), static_cast<uint32_t>((i >> 1) & 1), 64); };
      auto issue_q = [&](int upto) {
        while (q_issued <= upto && q_issued < n) {
          const int i = q_issued;
          if (i >= 2) wait_s(i - 2);               // S(i-2) has consumed this group's query tile
          const int t = t_begin + i, head = t / QT, qt = t - head * QT;
          const int b = head / a.H, h = head - b * a.H;
          mbar_arrive_expect_tx(bar(FB_Q + (i & 1)), kChunkBytes);
          tma_load_3d(&tmQ, bar(FB_Q + (i & 1)), sQ_u + (i & 1) * kChunkBytes, h * DH, qt * 128, b);
          ++q_issued;
        }
      };
      for (int i = 0; i < n; ++i) {
        const int t = t_begin + i, head = t / QT, qt = t - head * QT;
        if (i == 0 || qt == 0) {                   // first item of a head: its K and V replace the previous head's
          const int b = head / a.H, h = head - b * a.H;
          if (i > 0) wait_s(i - 1);                // every S MMA of the previous head has retired (in-order pipe)
          mbar_arrive_expect_tx(bar(FB_K), kv_bytes);
          tma_load_3d(&tmK, bar(FB_K), sK_u, h * DH, 0, b);
          issue_q(i + 1);
          if (i > 0) wait_o(i - 1);                // every P V MMA of the previous head has retired
          mbar_arrive_expect_tx(bar(FB_V), kv_bytes);
          tma_load_3d(&tmV, bar(FB_V), sV_u, h * DH, 0, b);
        } else {
          issue_q(i + 1);
        }
      }
    }
  } else if (warp == 17) {
    // =============================== tcgen05 issuer ===============================
    // Two independent cursors — the next O = P V and the next S = Q K^T — each issued as soon as ITS operands are
    // there (non-blocking probes).  With up to 224 keys the two S buffers (2 x 224 columns) leave room for one shared O
    // accumulator (64 columns): S(i+2) then only needs P(i) written (S(i) consumed), not O(i) drained, so it is in
    // flight while the group still waits for and drains O(i).  Wider rows fall back to O overwriting S's first columns.
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, NK, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(128, DH, false, true);
      const int nks = NK >> 4;
      const bool split = a.o_col >= 0;
      const uint64_t dK = desc_kmajor(sK_u), dV = desc_mnmajor(sV_u, 8192);
      uint32_t k_heads = 0, v_heads = 0;           // heads whose K / V have been consumed
      int s_next = 0, pv_next = 0;
      bool s_k = false, s_q = false, s_d = false, p_v = false, p_p = false, p_o = false;
      long long t_idle = clock64();
      while (pv_next < n) {
        bool did = false;
        if (pv_next < s_next) {                            // O(j) = P(j) V
          const int j = pv_next, t = t_begin + j, head = t / QT, qt = t - head * QT, g = j & 1;
          const bool newh = (j == 0 || qt == 0);
          if (!p_v) p_v = !newh || mbar_try_wait(bar(FB_V), v_heads & 1u);
          if (!p_p) p_p = mbar_try_wait(bar(FB_P + g), static_cast<uint32_t>((j >> 1) & 1));
          if (!p_o) p_o = !split || j == 0 || mbar_try_wait(bar(FB_D + ((j - 1) & 1)), static_cast<uint32_t>(((j - 1) >> 1) & 1));
          if (p_v && p_p && p_o) {
            if (newh) ++v_heads;
            tc_fence_after();
            const uint32_t d = tmem_base + static_cast<uint32_t>(split ? a.o_col : g * a.s_stride);
            const uint64_t dP = desc_kmajor(sP_u + g * p_bytes);
            for (int t16 = 0; t16 < nks; ++t16)    // 16 keys per step: 32 B inside a 64-key chunk, 16 KiB between chunks
              umma_bf16_ss(d, dP + static_cast<uint64_t>((t16 >> 2) * (kChunkBytes >> 4) + (t16 & 3) * 2), dV + static_cast<uint64_t>(t16 * 128),
                           idesc_o, t16 > 0 ? 1u : 0u);
            umma_commit(bar(FB_O + g));
            ++pv_next;
            p_v = p_p = p_o = false;
            did = true;
          }
        }
        if (s_next < n && (split || s_next <= pv_next + 1)) {   // S(i) into group (i & 1)'s buffer
          const int i = s_next, t = t_begin + i, head = t / QT, qt = t - head * QT, g = i & 1;
          const bool newh = (i == 0 || qt == 0);
          if (!s_k) s_k = !newh || mbar_try_wait(bar(FB_K), k_heads & 1u);
          if (!s_q) s_q = mbar_try_wait(bar(FB_Q + g), static_cast<uint32_t>((i >> 1) & 1));
          // the buffer is free once S(i-2) has been read (split: P(i-2) written) / once O(i-2), which overwrote it, is drained
          if (!s_d) s_d = i < 2 || mbar_try_wait(bar((split ? FB_P : FB_D) + g), static_cast<uint32_t>(((i - 2) >> 1) & 1));
          if (s_k && s_q && s_d) {
            if (newh) ++k_heads;
            tc_fence_after();
            const uint32_t d = tmem_base + static_cast<uint32_t>(g * a.s_stride);
            const uint64_t dQ = desc_kmajor(sQ_u + g * kChunkBytes);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(d, dQ + 2 * k, dK + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(bar(FB_S + g));
            ++s_next;
            s_k = s_q = s_d = false;
            did = true;
          }
        }
        if (did) {
          t_idle = clock64();
        } else {
          __nanosleep(20);
          if (clock64() - t_idle > 4000000000ll) __trap();   // a protocol bug must trap, never hang the GPU box
        }
      }
    }
  } else {
    // =============================== softmax groups ===============================
    const int g = warp >> 3;                       // group 0: items 0, 2, 4, ...; group 1: items 1, 3, 5, ...
    const int half = (warp >> 2) & 1;              // two threads per query row: which half of the keys / head-dim columns
    const int r = (warp & 3) * 32 + lane;          // query row within the tile == TMEM lane
    const int rx = r & 7;
    const bool elected = (r == 0 && half == 0);
    const uint32_t tquad = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t trow = tmem_base + static_cast<uint32_t>(g * a.s_stride) + tquad;
    const uint32_t torow = a.o_col >= 0 ? tmem_base + static_cast<uint32_t>(a.o_col) + tquad : trow;   // where O(i) lands
    const uint32_t sPg = sP_u + g * p_bytes;
    float* red_max = red + g * 512;                // [2 halves][128]
    float* red_sum = red_max + 256;
    // this thread's key columns [cb, ce): the 16-key units are split between the two halves of a row
    const int nunits = NK >> 4, umid = (nunits + 1) >> 1;
    const int cb = half ? umid * 16 : 0, ce = half ? NK : umid * 16;
    const uint64_t sc2 = pk2(a.scale_log2);
    // additive mask of the last 16-key unit (0 for keys < N, -inf for padding); mask_c0 < 0: no padding
    const int mask_c0 = (a.N < NK) ? NK - 16 : -1;
    uint64_t am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      am[j] = pk2((NK - 16 + 2 * j < a.N) ? 0.f : -INFINITY, (NK - 16 + 2 * j + 1 < a.N) ? 0.f : -INFINITY);
    for (int i = g; i < n; i += 2) {
      const uint32_t ph = static_cast<uint32_t>((i >> 1) & 1);
      const int t = t_begin + i, head = t / QT, qt = t - head * QT;
      const int b = head / a.H, h = head - b * a.H;
      mbar_wait_backoff(bar(FB_S + g), ph, 20);
      tc_fence_after();
      // pass 1: row max over this thread's keys
      float mx = -INFINITY;
      for (int c = cb; c < ce; c += 32) {
        uint32_t v[32];
        if (c + 32 <= ce) {
          tmem_ld_32x32b_x32(trow + c, v);
          tmem_ld_wait();
          if (c + 16 == mask_c0) add_mask16(v + 16, am);
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
          tmem_ld_32x32b_x16(trow + c, reinterpret_cast<uint32_t(&)[16]>(v));
          tmem_ld_wait();
          if (c == mask_c0) add_mask16(v, am);
#pragma unroll
          for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      red_max[half * 128 + r] = mx;
      // the TMA store of this group's previous O tile has finished reading the staging tile (chunk 0 of the P image)
      if (elected) bulk_wait_read<0>();
      named_bar_sync(1 + g, 256);
      mx = fmaxf(mx, red_max[(half ^ 1) * 128 + r]);
      // pass 2: p = exp2((s - max) * c), row sum, bf16 P into the swizzled A-operand image
      const uint64_t nm2 = pk2(-mx * a.scale_log2);
      uint64_t sum2a = pk2(0.f), sum2b = pk2(0.f);
      for (int c = cb; c < ce; c += 32) {
        uint32_t v[32];
        if (c + 32 <= ce) {
          tmem_ld_32x32b_x32(trow + c, v);
          tmem_ld_wait();
          if (c + 16 == mask_c0) add_mask16(v + 16, am);
          softmax_chunk<32>(v, sc2, nm2, sum2a, sum2b, sPg, r, rx, c);
        } else {
          tmem_ld_32x32b_x16(trow + c, reinterpret_cast<uint32_t(&)[16]>(v));
          tmem_ld_wait();
          if (c == mask_c0) add_mask16(v, am);
          softmax_chunk<16>(v, sc2, nm2, sum2a, sum2b, sPg, r, rx, c);
        }
      }
      float s0, s1, s2, s3;
      upk2(sum2a, s0, s1);
      upk2(sum2b, s2, s3);
      float sum = (s0 + s1) + (s2 + s3);
      red_sum[half * 128 + r] = sum;               // read by the other half after the P V barrier (release / acquire chain)
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(FB_P + g));
      // O = P V
      mbar_wait_backoff(bar(FB_O + g), ph, 20);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32b_x32(torow + static_cast<uint32_t>(half * 32), o);   // this thread's 32 of the 64 head-dim columns
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(FB_D + g));                  // O has been read: the accumulator may be overwritten
      sum += red_sum[(half ^ 1) * 128 + r];
      {                                            // O / sum -> bf16 -> staging tile (the P image is dead: P V has retired)
        const uint64_t inv2 = pk2(1.0f / sum);
        const uint32_t row_addr = sPg + static_cast<uint32_t>(r * 128);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float y0, y1;
            upk2(mul2(pk2(__uint_as_float(o[8 * u + 2 * j]), __uint_as_float(o[8 * u + 2 * j + 1])), inv2), y0, y1);
            w[j] = pack_bf16x2(y0, y1);
          }
          st_shared_v4(row_addr + static_cast<uint32_t>(((half * 4 + u) ^ rx) << 4), w[0], w[1], w[2], w[3]);
        }
      }
      const int row = qt * 128 + r;
      if (half == 0 && row < a.N && a.lse) a.lse[(static_cast<long long>(b) * a.H + h) * a.N + row] = mx * a.scale + logf(sum);
      fence_proxy_async_smem();
      named_bar_sync(1 + g, 256);
      if (elected) {                               // O tile [128 x 64] leaves as one TMA store (rows >= N clipped)
        tma_store_3d(&tmO, sPg, h * DH, qt * 128, b);
        bulk_commit();
      }
    }
    if (elected) bulk_wait_all();                  // shared memory must outlive the reads of the last store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ================================================================================================
// backward
// ================================================================================================
constexpr int kBwdThreads = 320;         // warps 0-7: P / dS / drains, warp 8: TMA producer, warp 9: tcgen05 issuer
// TMEM columns
constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256 /* + 64 qt */, T_DK = 384, T_DV = 448;

enum BwdBar {
  BB_KV = 0 /*[2] K,V stage loaded*/, BB_KVFREE = 2 /*[2]*/, BB_QDO = 4 /*[2] Q,dO stage loaded*/, BB_QFREE = 6 /*[2]*/,
  BB_O = 8 /* O tile loaded */, BB_OFREE = 9 /* 256: D_i computed */, BB_S = 10, BB_DP = 11, BB_P = 12 /*256*/, BB_DS = 13 /*256*/,
  BB_PFREE = 14 /* dV(i) retired: P image free */, BB_DSFREE = 15 /* dK(i), dQ(i) retired: dS image free, iteration complete */,
  BB_DVDR = 16 /*256: dV accumulator drained*/, BB_DKDR = 17 /*256*/, BB_DQDR = 18 /*[2] 256*/, BB_COUNT = 20
};

// One finished accumulator tile [128 x 64]: TMEM -> registers (the accumulator is released to the issuer) -> bf16 ->
// TMA store.  Every warp drains its own 32 rows x 32 columns through its own 2 KiB 64B-swizzled staging tile and its
// own bulk-store group: no CTA-wide barrier (the first versions synchronised all eight warps twice per drain, 12 times
// per head: 15 % of the warp samples sat there).  Inlined at three places only (see `drain` in the kernel).
__device__ __forceinline__ void drain_warp(uint32_t taddr, uint32_t drained_bar, const CUtensorMap* tm, uint32_t stage_w,
                                        int lane, int col, int row0, int b) {
  uint32_t v[32];
  tmem_ld_32x32b_x32(taddr, v);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(drained_bar);
  if (lane == 0) bulk_wait_read<0>();          // this warp's previous store has finished reading the staging tile
  __syncwarp();
  const uint32_t x = (static_cast<uint32_t>(lane) >> 1) & 3u;   // SWIZZLE_64B: 16-byte unit ^= address bits [7,9)
#pragma unroll
  for (int j = 0; j < 4; ++j)
    st_shared_v4(stage_w + static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(j) ^ x) << 4),
                 pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])),
                 pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                 pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                 pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(tm, stage_w, col, row0, b);   // rows >= N clipped by the tensor map
    bulk_commit();
  }
}

// exp2 of 2 x N packed scores in place (x <- exp2(x * c + nl)), optionally with the padding mask added to the 16 values
// at `mask_at`, then bf16 -> the 128B-swizzled image at key columns c0 .. c0 + W - 1 of row r
template <int W>
__device__ __forceinline__ void p_block(uint32_t* x, uint64_t sc2, uint64_t nl2, bool masked, int mask_at, const uint64_t (&am)[8],
                                        uint32_t image, uint32_t row_off, int rx, int c0) {
  if (masked) add_mask16(x + mask_at, am);
#pragma unroll
  for (int q = 0; q < W / 8; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x0, x1;
      upk2(fma2(pk2(__uint_as_float(x[8 * q + 2 * j]), __uint_as_float(x[8 * q + 2 * j + 1])), sc2, nl2), x0, x1);
      const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
      x[8 * q + 2 * j] = __float_as_uint(e0);
      x[8 * q + 2 * j + 1] = __float_as_uint(e1);
      w[j] = pack_bf16x2(e0, e1);
    }
    const int c = c0 + 8 * q;
    st_shared_v4(image + static_cast<uint32_t>((c >> 6) * kChunkBytes) + row_off + static_cast<uint32_t>((((c & 63) >> 3) ^ rx) << 4),
                 w[0], w[1], w[2], w[3]);
  }
}
// dS = P * (c dP - c D) -> bf16 -> image (the softmax scale c rides in here, so dQ / dK need no scaling when drained)
template <int W>
__device__ __forceinline__ void ds_block(const uint32_t* pv, const uint32_t* d, uint64_t c2, uint64_t ndc2, uint32_t image,
                                         uint32_t row_off, int rx, int c0) {
#pragma unroll
  for (int q = 0; q < W / 8; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t t2 = fma2(pk2(__uint_as_float(d[8 * q + 2 * j]), __uint_as_float(d[8 * q + 2 * j + 1])), c2, ndc2);
      float y0, y1;
      upk2(mul2(pk2(__uint_as_float(pv[8 * q + 2 * j]), __uint_as_float(pv[8 * q + 2 * j + 1])), t2), y0, y1);
      w[j] = pack_bf16x2(y0, y1);
    }
    const int c = c0 + 8 * q;
    st_shared_v4(image + static_cast<uint32_t>((c >> 6) * kChunkBytes) + row_off + static_cast<uint32_t>((((c & 63) >> 3) ^ rx) << 4),
                 w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_ws(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQ,
            const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
            const __grid_constant__ WsParams a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need 1024-byte aligned bases
  // shared memory: K,V stages [2][K | V] | Q,dO stages [2][Q | dO] | O tile | staging tile | P image (2 chunks) | dS image
  const uint32_t sKV_u = smem_u32(smem);
  const uint32_t sQD_u = sKV_u + 4 * kChunkBytes;
  const uint32_t sO_u = sQD_u + 4 * kChunkBytes;
  const uint32_t sStage_u = sO_u + kChunkBytes;
  const uint32_t sP_u = sStage_u + kChunkBytes;
  const uint32_t sDS_u = sP_u + 2 * kChunkBytes;
  uint8_t* tail = smem + 14 * kChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BB_COUNT);
  float* red = reinterpret_cast<float*>(bars + BB_COUNT + 2);   // [2 halves][128 rows] partial D_i
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * static_cast<uint32_t>(i); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmDQ); tma_prefetch_desc(&tmDK); tma_prefetch_desc(&tmDV);
    for (int i = 0; i < BB_COUNT; ++i) {
      const bool wide = (i == BB_OFREE || i == BB_P || i == BB_DS || i == BB_DVDR || i == BB_DKDR || i == BB_DQDR || i == BB_DQDR + 1);
      mbar_init(bar(i), wide ? 256 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int QT = a.qtiles, KT = a.ktiles, per_head = QT * KT;
  const int h_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * a.total / gridDim.x);
  const int h_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * a.total / gridDim.x);
  const int nheads = h_end - h_begin;
  const int n = nheads * per_head;                 // iterations of this CTA; iteration i = (hl, kt, qt), qt fastest

  if (warp == 8) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int hl = 0, kt = 0, qt = 0, b = h_begin / a.H, h = h_begin - b * a.H;
      for (int i = 0; i < n; ++i) {
        if (qt == 0) {                             // a new key tile: K, V into stage kc & 1
          const int kc = hl * KT + kt, st = kc & 1;
          if (kc >= 2) mbar_wait_backoff(bar(BB_KVFREE + st), static_cast<uint32_t>(((kc >> 1) - 1) & 1), 64);
          mbar_arrive_expect_tx(bar(BB_KV + st), 2 * kChunkBytes);
          tma_load_3d(&tmK, bar(BB_KV + st), sKV_u + st * 2 * kChunkBytes, h * DH, kt * 128, b);
          tma_load_3d(&tmV, bar(BB_KV + st), sKV_u + st * 2 * kChunkBytes + kChunkBytes, h * DH, kt * 128, b);
        }
        if (kt == 0) {                             // first use of this query tile: Q, dO into stage qc & 1, O into its tile
          const int qc = hl * QT + qt, st = qc & 1;
          if (qc >= 2) mbar_wait_backoff(bar(BB_QFREE + st), static_cast<uint32_t>(((qc >> 1) - 1) & 1), 64);
          mbar_arrive_expect_tx(bar(BB_QDO + st), 2 * kChunkBytes);
          tma_load_3d(&tmQ, bar(BB_QDO + st), sQD_u + st * 2 * kChunkBytes, h * DH, qt * 128, b);
          tma_load_3d(&tmDO, bar(BB_QDO + st), sQD_u + st * 2 * kChunkBytes + kChunkBytes, h * DH, qt * 128, b);
          if (qc >= 1) mbar_wait_backoff(bar(BB_OFREE), static_cast<uint32_t>((qc - 1) & 1), 64);
          mbar_arrive_expect_tx(bar(BB_O), kChunkBytes);
          tma_load_3d(&tmO, bar(BB_O), sO_u, h * DH, qt * 128, b);
        }
        if (++qt == QT) { qt = 0; if (++kt == KT) { kt = 0; ++hl; if (++h == a.H) { h = 0; ++b; } } }
      }
    }
  } else if (warp == 9) {
    // =============================== tcgen05 issuer ===============================
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, DH, true, true);     // dV, dK: A = P^T / dS^T (MN-major), B MN-major
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, DH, false, true);   // dQ: A = dS (K-major), B = K (MN-major)
      const int nkp_last = ((a.N - (KT - 1) * 128) + 15) & ~15;              // padded key count of the last key tile
      const uint32_t idesc_full = umma_idesc_bf16(128, 128, false, false), idesc_last = umma_idesc_bf16(128, nkp_last, false, false);
      // descriptors of the fixed buffers
      const uint64_t dPt = desc_mnmajor(sP_u, kChunkBytes), dDSt = desc_mnmajor(sDS_u, kChunkBytes), dDSk = desc_kmajor(sDS_u);
      // stage-dependent operands: K-major and MN-major views of the same tiles (stage stride = 2 tiles)
      const uint64_t dK0 = desc_kmajor(sKV_u), dKt0 = desc_mnmajor(sKV_u, 8192), dV0 = desc_kmajor(sKV_u + kChunkBytes);
      const uint64_t dQ0 = desc_kmajor(sQD_u), dQt0 = desc_mnmajor(sQD_u, 8192);
      const uint64_t dDO0 = desc_kmajor(sQD_u + kChunkBytes), dDOt0 = desc_mnmajor(sQD_u + kChunkBytes, 8192);
      constexpr uint64_t kStage = (2 * kChunkBytes) >> 4;
      // the coordinates of the iteration whose S / dP are issued next
      int s_hl = 0, s_kt = 0, s_qt = 0;
      auto issue_s_dp_wait = [&]() {               // operands of iteration (s_hl, s_kt, s_qt) have landed
        const int kc = s_hl * KT + s_kt, qc = s_hl * QT + s_qt;
        if (s_qt == 0) mbar_wait_backoff(bar(BB_KV + (kc & 1)), static_cast<uint32_t>((kc >> 1) & 1), 32);
        if (s_kt == 0) mbar_wait_backoff(bar(BB_QDO + (qc & 1)), static_cast<uint32_t>((qc >> 1) & 1), 32);
        tc_fence_after();
      };
      auto issue_s = [&]() {                       // S = Q K^T (N = this key tile's padded key count)
        const int kc = s_hl * KT + s_kt, qc = s_hl * QT + s_qt;
        const uint64_t dQ = dQ0 + (qc & 1) * kStage, dK = dK0 + (kc & 1) * kStage;
        const uint32_t idesc = (s_kt == KT - 1) ? idesc_last : idesc_full;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + T_S, dQ + 2 * k, dK + 2 * k, idesc, k > 0 ? 1u : 0u);
        umma_commit(bar(BB_S));
      };
      auto issue_dp = [&]() {                      // dP = dO V^T
        const int kc = s_hl * KT + s_kt, qc = s_hl * QT + s_qt;
        const uint64_t dDO = dDO0 + (qc & 1) * kStage, dV = dV0 + (kc & 1) * kStage;
        const uint32_t idesc = (s_kt == KT - 1) ? idesc_last : idesc_full;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + T_DP, dDO + 2 * k, dV + 2 * k, idesc, k > 0 ? 1u : 0u);
        umma_commit(bar(BB_DP));
      };
      auto advance_s = [&]() { if (++s_qt == QT) { s_qt = 0; if (++s_kt == KT) { s_kt = 0; ++s_hl; } } };
      issue_s_dp_wait();
      issue_s();
      issue_dp();
      int hl = 0, kt = 0, qt = 0;
      for (int i = 0; i < n; ++i) {
        const int kc = hl * KT + kt, qc = hl * QT + qt;
        const uint32_t ph = static_cast<uint32_t>(i & 1);
        const uint64_t dKt = dKt0 + (kc & 1) * kStage, dQt = dQt0 + (qc & 1) * kStage, dDOt = dDOt0 + (qc & 1) * kStage;
        const int nks = (kt == KT - 1 ? nkp_last : 128) >> 4;
        advance_s();                               // (s_*) = iteration i + 1
        // ---- P(i) is in shared memory (and S has been read): S(i+1), then dV += P^T dO
        mbar_wait_backoff(bar(BB_P), ph, 20);
        tc_fence_after();
        if (i + 1 < n) { issue_s_dp_wait(); issue_s(); }
        if (qt == 0 && kc > 0) { mbar_wait_backoff(bar(BB_DVDR), static_cast<uint32_t>((kc - 1) & 1), 20); tc_fence_after(); }
#pragma unroll
        for (int k = 0; k < 8; ++k)        // K = 128 query rows; A = P^T (MN-major image of sP), B = dO (MN-major)
          umma_bf16_ss(tmem_base + T_DV, dPt + 128 * k, dDOt + 128 * k, idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar(BB_PFREE));
        // ---- dS(i) is in shared memory (and dP has been read): dP(i+1), then dK += dS^T Q, dQ += dS K
        mbar_wait_backoff(bar(BB_DS), ph, 20);
        tc_fence_after();
        if (i + 1 < n) issue_dp();
        if (qt == 0 && kc > 0) { mbar_wait_backoff(bar(BB_DKDR), static_cast<uint32_t>((kc - 1) & 1), 20); tc_fence_after(); }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tmem_base + T_DK, dDSt + 128 * k, dQt + 128 * k, idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
        if (kt == 0 && hl > 0) { mbar_wait_backoff(bar(BB_DQDR + qt), static_cast<uint32_t>((hl - 1) & 1), 20); tc_fence_after(); }
        for (int t16 = 0; t16 < nks; ++t16)   // K = this tile's keys; A = dS (K-major over keys), B = K (MN-major: keys x dh)
          umma_bf16_ss(tmem_base + T_DQ + static_cast<uint32_t>(qt * 64),
                       dDSk + static_cast<uint64_t>((t16 >> 2) * (kChunkBytes >> 4) + (t16 & 3) * 2), dKt + static_cast<uint64_t>(t16 * 128),
                       idesc_dq, (kt > 0 || t16 > 0) ? 1u : 0u);
        umma_commit(bar(BB_DSFREE));
        if (qt == QT - 1) umma_commit(bar(BB_KVFREE + (kc & 1)));   // last user of this K, V stage
        if (kt == KT - 1) umma_commit(bar(BB_QFREE + (qc & 1)));    // last user of this Q, dO stage
        if (++qt == QT) { qt = 0; if (++kt == KT) { kt = 0; ++hl; } }
      }
    }
  } else {
    // =============================== CUDA-core warps: P, dS, drains ===============================
    const int half = warp >> 2;                    // two threads per query row; `half` picks the keys
    const int r = (warp & 3) * 32 + lane;          // query row within the tile == TMEM lane
    const int rx = r & 7;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t row_off = static_cast<uint32_t>(r * 128);
    const uint32_t stage_w = sStage_u + static_cast<uint32_t>(warp * 2048);   // this warp's 32 x 32 bf16 store tile
    const int drain_col = half * 32, drain_row = (warp & 3) * 32;             // its corner inside a [128 x 64] tile
    float lse2_q0 = 0.f, lse2_q1 = 0.f, d_q0 = 0.f, d_q1 = 0.f;   // per query tile: LSE * log2(e) and c * D_i of this thread's row
    // accumulators waiting to be drained: (image, head) and tile they belong to
    bool pend_dv = false, pend_dk = false, pend_dq0 = false, pend_dq1 = false;
    int pend_kv_b = 0, pend_kv_h = 0, pend_kv_kt = 0, pend_dq0_b = 0, pend_dq0_h = 0, pend_dq1_b = 0, pend_dq1_h = 0;
    const uint64_t sc2 = pk2(a.scale_log2), c2 = pk2(a.scale);
    // additive mask of the last 16-key unit of the last key tile (0 for keys < N, -inf for padding)
    const int last_valid = a.N - (KT - 1) * 128;                  // keys of the last key tile (1 .. 128)
    const int last_unit0 = ((last_valid + 15) & ~15) - 16;        // first key of its last unit
    uint64_t am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      am[j] = pk2((last_unit0 + 2 * j < last_valid) ? 0.f : -INFINITY, (last_unit0 + 2 * j + 1 < last_valid) ? 0.f : -INFINITY);

    // accumulator w: 0 dV, 1 dK, 2 dQ of query tile 0, 3 dQ of query tile 1 (their "drained" barriers are consecutive)
    auto drain = [&](int w) {
      const uint32_t tcol = w == 0 ? T_DV : (w == 1 ? T_DK : T_DQ + static_cast<uint32_t>((w - 2) * 64));
      const CUtensorMap* tm = w == 0 ? &tmDV : (w == 1 ? &tmDK : &tmDQ);
      const int hh = w < 2 ? pend_kv_h : (w == 2 ? pend_dq0_h : pend_dq1_h);
      const int bb = w < 2 ? pend_kv_b : (w == 2 ? pend_dq0_b : pend_dq1_b);
      const int row0 = w < 2 ? pend_kv_kt * 128 : (w - 2) * 128;
      drain_warp(trow + tcol + static_cast<uint32_t>(drain_col), bar(BB_DVDR + w), tm, stage_w, lane, hh * DH + drain_col,
                 row0 + drain_row, bb);
      if (w == 0) pend_dv = false;
      else if (w == 1) pend_dk = false;
      else if (w == 2) pend_dq0 = false;
      else pend_dq1 = false;
    };

    int hl = 0, kt = 0, qt = 0, b = h_begin / a.H, h = h_begin - b * a.H;
    // LSE of this thread's row for the next iteration that starts a query tile: requested one iteration ahead and only
    // touched when that iteration begins, so the load never stalls the pipeline
    float lse_carry = INFINITY;
    if (n > 0 && r < a.N) lse_carry = a.lse[(static_cast<long long>(b) * a.H + h) * a.N + r];
    for (int i = 0; i < n; ++i) {
      const int qc = hl * QT + qt;
      const uint32_t ph = static_cast<uint32_t>(i & 1);
      const float lse_cur = lse_carry;
      // coordinates of the next iteration; its LSE is requested now if it starts a query tile
      int nqt = qt + 1, nkt = kt, nhl = hl, nb = b, nh = h;
      if (nqt == QT) { nqt = 0; if (++nkt == KT) { nkt = 0; ++nhl; if (++nh == a.H) { nh = 0; ++nb; } } }
      if (i + 1 < n && nkt == 0) {
        const int row = nqt * 128 + r;
        lse_carry = (row < a.N) ? a.lse[(static_cast<long long>(nb) * a.H + nh) * a.N + row] : INFINITY;
      }
      if (kt == 0) {
        // D_i = rowsum(dO * O) of this query tile from the swizzled tiles
        mbar_wait_backoff(bar(BB_QDO + (qc & 1)), static_cast<uint32_t>((qc >> 1) & 1), 20);
        mbar_wait_backoff(bar(BB_O), static_cast<uint32_t>(qc & 1), 20);
        const uint32_t sDO = sQD_u + (qc & 1) * 2 * kChunkBytes + kChunkBytes;
        float part = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t off = row_off + static_cast<uint32_t>(((half * 4 + u) ^ rx) << 4);
          part += dot8_bf16(ld_shared_v4(sO_u + off), ld_shared_v4(sDO + off));
        }
        red[half * 128 + r] = part;
        named_bar_sync(1, 256);
        const float di = (red[r] + red[128 + r]) * a.scale;
        mbar_arrive(bar(BB_OFREE));                // the O tile may be replaced
        named_bar_sync(1, 256);                    // red[] may be rewritten by the next query tile
        if (qt == 0) { d_q0 = di; lse2_q0 = lse_cur * kLog2e; } else { d_q1 = di; lse2_q1 = lse_cur * kLog2e; }
      }
      const float lse2 = qt == 0 ? lse2_q0 : lse2_q1;
      const float Dc = qt == 0 ? d_q0 : d_q1;
      // this thread's keys: 32-key pairs go to the halves in turn (pair p -> half p & 1: slot 0 = pair `half`, slot 1 = pair
      // `half + 2`), an odd last 16-key unit to the half with fewer pairs (its slot 1); only the tile's last unit is masked
      const int nk_valid = min(128, a.N - kt * 128);
      const int nunits = (nk_valid + 15) >> 4, npairs = nunits >> 1;
      const bool part_last = (nk_valid & 15) != 0;
      const bool s0 = half < npairs;
      const bool s1p = half + 2 < npairs;
      const bool s1u = !s1p && (nunits & 1) && ((npairs & 1) == half);
      const int c_s0 = 32 * half, c_s1 = s1p ? 64 + 32 * half : 16 * (nunits - 1);
      const bool m0 = s0 && part_last && (2 * half + 1 == nunits - 1);        // the mask applies to the pair's second unit
      const bool m1p = s1p && part_last && (2 * (half + 2) + 1 == nunits - 1);
      const bool m1u = s1u && part_last;
      const uint64_t nl2 = pk2(-lse2);
      uint32_t p[64];                                         // P of this thread's keys (fp32), kept for dS
      // ---- P = exp2(S*c - LSE*log2e) -> bf16 -> sP  (rows >= N: LSE = +inf -> 0; padding keys -> 0)
      mbar_wait_backoff(bar(BB_S), ph, 20);
      if (i >= 1) mbar_wait_backoff(bar(BB_PFREE), static_cast<uint32_t>((i - 1) & 1), 20);   // dV(i-1) has read the previous P
      tc_fence_after();
      if (s0) tmem_ld_32x32b_x32(trow + T_S + static_cast<uint32_t>(c_s0), reinterpret_cast<uint32_t(&)[32]>(p[0]));
      if (s1p) tmem_ld_32x32b_x32(trow + T_S + static_cast<uint32_t>(c_s1), reinterpret_cast<uint32_t(&)[32]>(p[32]));
      else if (s1u) tmem_ld_32x32b_x16(trow + T_S + static_cast<uint32_t>(c_s1), reinterpret_cast<uint32_t(&)[16]>(p[32]));
      tmem_ld_wait();
      if (s0) p_block<32>(p, sc2, nl2, m0, 16, am, sP_u, row_off, rx, c_s0);
      if (s1p) p_block<32>(p + 32, sc2, nl2, m1p, 16, am, sP_u, row_off, rx, c_s1);
      else if (s1u) p_block<16>(p + 32, sc2, nl2, m1u, 0, am, sP_u, row_off, rx, c_s1);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(BB_P));
      // ---- slot after P: the MMAs of iteration i-1 have retired by now; drain what they completed
      if (i >= 1) { mbar_wait_backoff(bar(BB_DSFREE), static_cast<uint32_t>((i - 1) & 1), 20); tc_fence_after(); }
      {
        const int w = pend_dv ? 0 : ((qt == 1 && pend_dq0) ? 2 : ((qt == 0 && pend_dq1) ? 3 : -1));
        if (w >= 0) drain(w);
      }
      // ---- dS = P * c (dP - D) -> bf16 -> sdS
      mbar_wait_backoff(bar(BB_DP), ph, 20);
      tc_fence_after();
      {
        // one slot at a time: P (64 registers) stays live, and the register file of a 320-thread CTA ends at 168
        const uint64_t ndc2 = pk2(-Dc);
        uint32_t d[32];
        if (s0) {
          tmem_ld_32x32b_x32(trow + T_DP + static_cast<uint32_t>(c_s0), d);
          tmem_ld_wait();
          ds_block<32>(p, d, c2, ndc2, sDS_u, row_off, rx, c_s0);
        }
        if (s1p) {
          tmem_ld_32x32b_x32(trow + T_DP + static_cast<uint32_t>(c_s1), d);
          tmem_ld_wait();
          ds_block<32>(p + 32, d, c2, ndc2, sDS_u, row_off, rx, c_s1);
        } else if (s1u) {
          tmem_ld_32x32b_x16(trow + T_DP + static_cast<uint32_t>(c_s1), reinterpret_cast<uint32_t(&)[16]>(d));
          tmem_ld_wait();
          ds_block<16>(p + 32, d, c2, ndc2, sDS_u, row_off, rx, c_s1);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(BB_DS));
      // ---- slot after dS
#pragma unroll 1
      for (int k = 0; k < 2; ++k) {
        const int w = k == 0 ? (pend_dk ? 1 : -1) : ((qt == 0 && pend_dq0) ? 2 : ((qt == 1 && pend_dq1) ? 3 : -1));
        if (w >= 0) drain(w);
      }
      // what this iteration completes (drained one slot later, once its MMAs have retired)
      if (qt == QT - 1) { pend_dv = pend_dk = true; pend_kv_b = b; pend_kv_h = h; pend_kv_kt = kt; }
      if (kt == KT - 1) {
        if (qt == 0) { pend_dq0 = true; pend_dq0_b = b; pend_dq0_h = h; }
        else { pend_dq1 = true; pend_dq1_b = b; pend_dq1_h = h; }
      }
      qt = nqt; kt = nkt; hl = nhl; b = nb; h = nh;
    }
    if (n > 0) {
      mbar_wait_backoff(bar(BB_DSFREE), static_cast<uint32_t>((n - 1) & 1), 20);
      tc_fence_after();
#pragma unroll 1
      for (int w = 0; w < 4; ++w)
        if (w == 0 ? pend_dv : (w == 1 ? pend_dk : (w == 2 ? pend_dq0 : pend_dq1))) drain(w);
    }
    if (lane == 0) bulk_wait_all();                // shared memory must outlive the reads of this warp's last store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ================================================================================================
// host
// ================================================================================================
int head_map(CUtensorMap* m, const void* base, int H, int N, int B, long long row_stride, long long batch_stride, int box_rows) {
  uint64_t dims[3] = {(uint64_t)H * DH, (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
  uint32_t box[3] = {DH, (uint32_t)box_rows, 1};
  return vitb_make_tmap_nd_bf16(m, base, 3, dims, str, box);
}

// per-warp gradient stores of the backward: 32 rows x 32 head-dim columns, SWIZZLE_64B
int store_map32(CUtensorMap* m, const void* base, int H, int N, int B, long long row_stride, long long batch_stride) {
  uint64_t dims[3] = {(uint64_t)H * DH, (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
  uint32_t box[3] = {32, 32, 1};
  return vitb_make_tmap_nd_bf16_sw64(m, base, 3, dims, str, box);
}

int fwd_smem_bytes(int N) {
  const int NK = (N + 15) & ~15;
  const int kv_bytes = NK * 128, pchunks = (NK + 63) / 64;
  // K | V | 2 query tiles | 2 P images | barriers + TMEM slot | row max / sum exchange
  return 2 * kv_bytes + 2 * kChunkBytes + 2 * pchunks * kChunkBytes + 8 * (FB_COUNT + 2) + 2 * 512 * 4;
}

int check_ws(const vitb_attn_params* p, const char* who) {
  VITB_REQUIRE(p && p->struct_bytes == (int)sizeof(vitb_attn_params), VITB_ERR_BAD_ARG, "%s: ABI mismatch", who);
  VITB_REQUIRE(p->dtype == VITB_BF16, VITB_ERR_UNSUPPORTED_SHAPE, "%s: bf16 only", who);
  VITB_REQUIRE(p->head_dim == DH, VITB_ERR_UNSUPPORTED_SHAPE, "%s: head_dim %d (only 64)", who, p->head_dim);
  VITB_REQUIRE(p->Nq == p->Nk && p->Nk >= 1 && p->Nk <= 256, VITB_ERR_UNSUPPORTED_SHAPE,
               "%s: Nq=%d Nk=%d (need Nq == Nk <= 256)", who, p->Nq, p->Nk);
  VITB_REQUIRE(p->q && p->k && p->v && p->o, VITB_ERR_BAD_ARG, "%s: null tensor", who);
  VITB_REQUIRE(p->o_row_stride % 8 == 0 && p->o_batch_stride % 8 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "%s: o strides %% 8", who);
  return VITB_OK;
}

}  // namespace

/* which: 0 = forward, 1 = backward.  The forward keeps two P images next to K, V and the query tiles: more than 240
 * keys do not fit the 227 KiB of shared memory (such shapes stay on vitb_attn_fwd_tc). */
extern "C" int vitb_attn_ws_supported(int which, int head_dim, int Nq, int Nk) {
  if (!(head_dim == DH && Nq == Nk && Nk >= 1 && Nk <= 256)) return 0;
  if (which == 0) return fwd_smem_bytes(Nk) <= 227 * 1024;
  return 1;
}

extern "C" int vitb_attn_fwd_ws(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  st = check_ws(p, "attn_fwd_ws");
  if (st != VITB_OK) return st;
  if (p->B == 0) return VITB_OK;
  const int N = p->Nk, NK = (N + 15) & ~15;
  const int smem = fwd_smem_bytes(N);
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_fwd_ws: %d keys need %d B of shared memory", N, smem);
  CUtensorMap tq, tk, tv, to;
  if ((st = head_map(&tq, p->q, p->H, N, p->B, p->q_row_stride, p->q_batch_stride, 128)) != VITB_OK) return st;
  if ((st = head_map(&tk, p->k, p->H, N, p->B, p->k_row_stride, p->k_batch_stride, NK)) != VITB_OK) return st;
  if ((st = head_map(&tv, p->v, p->H, N, p->B, p->v_row_stride, p->v_batch_stride, NK)) != VITB_OK) return st;
  if ((st = head_map(&to, p->o, p->H, N, p->B, p->o_row_stride, p->o_batch_stride, 128)) != VITB_OK) return st;
  WsParams a{};
  a.N = N; a.NK = NK; a.H = p->H;
  a.qtiles = (N + 127) / 128;
  a.ktiles = a.qtiles;
  a.total = p->B * p->H * a.qtiles;
  a.scale = 1.0f / sqrtf((float)DH);
  a.scale_log2 = a.scale * kLog2e;
  a.lse = p->lse;
  // TMEM: S buffers at columns 0 / 256, O(i) overwrites the first 64 columns of S(i).  (A layout with S at 0 / 224 and one
  // shared O accumulator at 448, which lets S(i+2) start before O(i) is drained, measured 6 % SLOWER at N = 197 —
  // profiles/attn_ws_r02.txt; the kernel still understands it: s_stride = 224, o_col = 448.)
  a.s_stride = 256;
  a.o_col = -1;
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int sms = vitb_num_sms();
  const int grid = a.total < sms ? a.total : sms;
  VITB_CUDA_CHECK(vitb_launch<kPdlAttn>(attn_fwd_ws, dim3(grid), dim3(kFwdThreads), smem, reinterpret_cast<cudaStream_t>(stream_), tq, tk,
                              tv, to, a));
  VITB_LAUNCH_CHECK("attn_fwd_ws");
  return VITB_OK;
}

extern "C" int vitb_attn_bwd_ws(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  st = check_ws(p, "attn_bwd_ws");
  if (st != VITB_OK) return st;
  if (p->B == 0) return VITB_OK;
  VITB_REQUIRE(p->lse && p->dout && p->dq && p->dk && p->dv, VITB_ERR_BAD_ARG, "attn_bwd_ws: null tensor");
  VITB_REQUIRE(p->dq_row_stride % 8 == 0 && p->dk_row_stride % 8 == 0 && p->dv_row_stride % 8 == 0 &&
                   p->do_row_stride % 8 == 0 && p->dq_batch_stride % 8 == 0 && p->dk_batch_stride % 8 == 0 &&
                   p->dv_batch_stride % 8 == 0 && p->do_batch_stride % 8 == 0,
               VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_ws: gradient strides %% 8");
  const int N = p->Nk;
  CUtensorMap tq, tk, tv, tdo, to, tdq, tdk, tdv;
  if ((st = head_map(&tq, p->q, p->H, N, p->B, p->q_row_stride, p->q_batch_stride, 128)) != VITB_OK) return st;
  if ((st = head_map(&tk, p->k, p->H, N, p->B, p->k_row_stride, p->k_batch_stride, 128)) != VITB_OK) return st;
  if ((st = head_map(&tv, p->v, p->H, N, p->B, p->v_row_stride, p->v_batch_stride, 128)) != VITB_OK) return st;
  if ((st = head_map(&tdo, p->dout, p->H, N, p->B, p->do_row_stride, p->do_batch_stride, 128)) != VITB_OK) return st;
  if ((st = head_map(&to, p->o, p->H, N, p->B, p->o_row_stride, p->o_batch_stride, 128)) != VITB_OK) return st;
  if ((st = store_map32(&tdq, p->dq, p->H, N, p->B, p->dq_row_stride, p->dq_batch_stride)) != VITB_OK) return st;
  if ((st = store_map32(&tdk, p->dk, p->H, N, p->B, p->dk_row_stride, p->dk_batch_stride)) != VITB_OK) return st;
  if ((st = store_map32(&tdv, p->dv, p->H, N, p->B, p->dv_row_stride, p->dv_batch_stride)) != VITB_OK) return st;
  WsParams a{};
  a.N = N; a.NK = (N + 15) & ~15; a.H = p->H;
  a.qtiles = (N + 127) / 128;
  a.ktiles = a.qtiles;
  a.total = p->B * p->H;
  a.scale = 1.0f / sqrtf((float)DH);
  a.scale_log2 = a.scale * kLog2e;
  a.lse = p->lse;
  // 14 tiles of 16 KiB | barriers + TMEM slot | D_i partials
  const int smem = 14 * kChunkBytes + 8 * (BB_COUNT + 2) + 2 * 128 * 4;
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_ws: %d B of shared memory", smem);
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int sms = vitb_num_sms();
  const int grid = a.total < sms ? a.total : sms;
  VITB_CUDA_CHECK(vitb_launch<kPdlAttn>(attn_bwd_ws, dim3(grid), dim3(kBwdThreads), smem, reinterpret_cast<cudaStream_t>(stream_), tq, tk,
                              tv, tdo, to, tdq, tdk, tdv, a));
  VITB_LAUNCH_CHECK("attn_bwd_ws");
  return VITB_OK;
}
