// vitb_attention_ws.cu — persistent, warp-specialised SelfAttention forward / backward on tcgen05 (bf16 in, fp32
// in TMEM) for head_dim 64 and up to 256 tokens (ViT-B/L at 224 px: N = 197 or 50).
//
// Replaces the q k^T / softmax / v chain of SelfAttention.forward (src/model.py:90-97) and Attention.forward
// (res-vit/model.py:273-293) plus its autograd backward — same contract as vitb_attn_fwd_tc / vitb_attn_bwd_tc.
//
// Why a second generation: the one-CTA-per-tile kernels of vitb_attention_tc.cu run every phase of their chain
// (TMA -> MMA -> CUDA-core pass -> CTA barrier -> MMA -> store) exposed; ncu showed 12-14 % tensor-pipe activity and
// 23-46 % issue utilisation (profiles/ncu_r01c.txt), i.e. latency-bound.  Here ONE CTA per SM stays resident and
// walks a contiguous range of work items; a TMA producer warp, a single-thread tcgen05 issuer warp and eight
// CUDA-core warps run as a pipeline connected by mbarriers, so loads, MMAs and the softmax arithmetic of
// neighbouring items overlap:
//
//   forward  (attn_fwd_ws)  item = (image, head, 128-query tile).  Two softmax groups of 128 threads (one thread per
//            query row, no cross-thread max / sum exchange) alternate over the items; each owns a 256-column TMEM
//            buffer (S, later O in its first 64 columns) and a P image in shared memory.  K / V are loaded once per
//            head and shared by its tiles.  While group g does the softmax of item i, the issuer has S(i+1) in flight
//            for the other group and P(i-1) V behind it.
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

constexpr int kWsThreads = 320;          // warps 0-7: CUDA-core work, warp 8: TMA producer, warp 9: tcgen05 issuer
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
  float scale;       // 1/sqrt(dh)
  float scale_log2;  // scale * log2(e)
  float* lse;     // [B,H,N]
};

// ================================================================================================
// forward
// ================================================================================================
// barriers (8 bytes each)
enum FwdBar { FB_K = 0, FB_V = 1, FB_Q = 2 /*[2]*/, FB_S = 4 /*[2]*/, FB_P = 6 /*[2]*/, FB_O = 8 /*[2]*/, FB_D = 10 /*[2]*/, FB_COUNT = 12 };

__global__ void __launch_bounds__(kWsThreads, 1)
attn_fwd_ws(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
            const __grid_constant__ WsParams a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int NK = a.NK;
  const int kv_bytes = NK * 128;                 // [NK keys x 64] bf16, 128B-swizzled rows
  const int pchunks = (NK + 63) >> 6;            // 64-key chunks of a P image
  const int p_bytes = pchunks * kChunkBytes;
  const uint32_t sK_u = smem_u32(smem);
  const uint32_t sV_u = sK_u + kv_bytes;
  const uint32_t sQ_u = sV_u + kv_bytes;         // [2] query tiles, one per softmax group
  const uint32_t sP_u = sQ_u + 2 * kChunkBytes;  // [2] P images; chunk 0 doubles as the O staging tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kv_bytes + 2 * kChunkBytes + 2 * p_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + FB_COUNT);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * static_cast<uint32_t>(i); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int i = 0; i < FB_COUNT; ++i) mbar_init(bar(i), (i >= FB_P && i < FB_O) || i >= FB_D ? 128 : 1);
    fence_barrier_init();
  }
  if (warp == 9) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
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

  if (warp == 8) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int q_issued = 0;
      auto wait_s = [&](int i) { mbar_wait(bar(FB_S + (i & 1)), static_cast<uint32_t>((i >> 1) & 1)); };
      auto wait_o = [&](int i) { mbar_wait(bar(FB_O + (i & 1)), static_cast<uint32_t>((i >> 1) & 1)); };
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
  } else if (warp == 9) {
    // =============================== tcgen05 issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, NK, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(128, DH, false, true);
      const int nks = NK >> 4;
      uint32_t k_heads = 0, v_heads = 0;           // heads whose K / V have been waited for
      for (int i = 0; i <= n; ++i) {
        if (i < n) {                               // S(i) = Q K^T into this group's TMEM buffer
          const int t = t_begin + i, head = t / QT, qt = t - head * QT, g = i & 1;
          if (i == 0 || qt == 0) { mbar_wait(bar(FB_K), k_heads & 1u); ++k_heads; }
          mbar_wait(bar(FB_Q + g), static_cast<uint32_t>((i >> 1) & 1));
          if (i >= 2) mbar_wait(bar(FB_D + g), static_cast<uint32_t>(((i - 2) >> 1) & 1));   // O(i-2) drained
          tc_fence_after();
          const uint32_t d = tmem_base + static_cast<uint32_t>(g * 256);
          const uint32_t q = sQ_u + g * kChunkBytes;
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16_ss(d, umma_smem_desc_sw128(q + k * 32, 16, 1024), umma_smem_desc_sw128(sK_u + k * 32, 16, 1024),
                         idesc_s, k > 0 ? 1u : 0u);
          umma_commit(bar(FB_S + g));
        }
        if (i >= 1) {                              // O(j) = P(j) V over the first 64 columns of S(j)'s buffer
          const int j = i - 1, t = t_begin + j, head = t / QT, qt = t - head * QT, g = j & 1;
          if (j == 0 || qt == 0) { mbar_wait(bar(FB_V), v_heads & 1u); ++v_heads; }
          mbar_wait(bar(FB_P + g), static_cast<uint32_t>((j >> 1) & 1));
          tc_fence_after();
          const uint32_t d = tmem_base + static_cast<uint32_t>(g * 256);
          const uint32_t p = sP_u + g * p_bytes;
          for (int t16 = 0; t16 < nks; ++t16)
            umma_bf16_ss(d, umma_smem_desc_sw128(p + (t16 >> 2) * kChunkBytes + (t16 & 3) * 32, 16, 1024),
                         umma_smem_desc_sw128(sV_u + t16 * 2048, 8192, 1024), idesc_o, t16 > 0 ? 1u : 0u);
          umma_commit(bar(FB_O + g));
        }
      }
    }
  } else {
    // =============================== softmax groups ===============================
    const int g = warp >> 2;                       // group 0: items 0, 2, 4, ...; group 1: items 1, 3, 5, ...
    const int r = (warp & 3) * 32 + lane;          // query row within the tile == TMEM lane
    const bool elected = (r == 0);
    const uint32_t trow = tmem_base + static_cast<uint32_t>(g * 256) + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t sPg = sP_u + g * p_bytes;
    const int nch = (NK + 31) >> 5;
    for (int i = g; i < n; i += 2) {
      const uint32_t ph = static_cast<uint32_t>((i >> 1) & 1);
      const int t = t_begin + i, head = t / QT, qt = t - head * QT;
      const int b = head / a.H, h = head - b * a.H;
      mbar_wait(bar(FB_S + g), ph);
      tc_fence_after();
      // pass 1: row max over the valid keys (two TMEM loads in flight per wait)
      float mx = -INFINITY;
      for (int c = 0; c < nch; c += 2) {
        uint32_t v0[32], v1[32];
        const bool two = (c + 1 < nch);
        issue_chunk(trow, c * 32, NK, v0);
        if (two) issue_chunk(trow, (c + 1) * 32, NK, v1);
        tmem_ld_wait();
        mx = fmaxf(mx, chunk_max(v0, c * 32, NK, a.N));
        if (two) mx = fmaxf(mx, chunk_max(v1, (c + 1) * 32, NK, a.N));
      }
      // the TMA store of this group's previous O tile has finished reading the staging tile (chunk 0 of the P image)
      if (i >= 2) {
        if (elected) bulk_wait_read<0>();
        named_bar_sync(1 + g, 128);
      }
      // pass 2: p = exp2((s - max) * c), row sum, bf16 P into the swizzled A-operand image
      const float mxs = mx * a.scale_log2;
      const uint64_t sc2 = pk2(a.scale_log2), nm2 = pk2(-mxs);
      uint64_t sum2a = pk2(0.f), sum2b = pk2(0.f);
      for (int c = 0; c < nch; ++c) {
        const int c0 = c * 32;
        uint32_t v[32];
        ld_chunk(trow, c0, NK, 0u, v);
        uint32_t pk[16];
        const bool interior = (c0 + 32 <= a.N);    // every column of the chunk is a valid key
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x0, x1;
          upk2(fma2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc2, nm2), x0, x1);
          float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
          if (!interior) {
            e0 = (c0 + 2 * j < a.N) ? e0 : 0.f;
            e1 = (c0 + 2 * j + 1 < a.N) ? e1 : 0.f;
          }
          const uint64_t e2 = pk2(e0, e1);
          if (j & 1) sum2b = add2(sum2b, e2); else sum2a = add2(sum2a, e2);
          pk[j] = pack_bf16x2(e0, e1);
        }
        const int kc = c0 >> 6, u0 = (c0 & 63) >> 3, nunits = (c0 + 32 <= NK) ? 4 : 2;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (u < nunits)
            st_shared_v4(sPg + kc * kChunkBytes + swz_unit(r, u0 + u), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(FB_P + g));
      float s0, s1, s2, s3;
      upk2(sum2a, s0, s1);
      upk2(sum2b, s2, s3);
      const float sum = (s0 + s1) + (s2 + s3);
      // O = P V
      mbar_wait(bar(FB_O + g), ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(trow, o0);
      tmem_ld_32x32b_x32(trow + 32u, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(FB_D + g));                  // the buffer may receive S(i+2)
      const float inv = 1.0f / sum;
      stage_row32_bf16(sPg, r, 0, o0, inv);        // the P image is dead: the P V MMAs have retired
      stage_row32_bf16(sPg, r, 1, o1, inv);
      const int row = qt * 128 + r;
      if (row < a.N && a.lse) a.lse[(static_cast<long long>(b) * a.H + h) * a.N + row] = mx * a.scale + logf(sum);
      fence_proxy_async_smem();
      named_bar_sync(1 + g, 128);
      if (elected) {                               // O tile [128 x 64] leaves as one TMA store (rows >= N clipped)
        tma_store_3d(&tmO, sPg, h * DH, qt * 128, b);
        bulk_commit();
      }
    }
    if (elected) bulk_wait_all();                  // shared memory must outlive the reads of the last store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ================================================================================================
// backward
// ================================================================================================
// TMEM columns
constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256 /* + 64 qt */, T_DK = 384, T_DV = 448;

enum BwdBar {
  BB_KV = 0 /*[2] K,V stage loaded*/, BB_KVFREE = 2 /*[2]*/, BB_QDO = 4 /*[2] Q,dO stage loaded*/, BB_QFREE = 6 /*[2]*/,
  BB_O = 8 /* O tile loaded */, BB_OFREE = 9 /* 256: D_i computed */, BB_S = 10, BB_DP = 11, BB_P = 12 /*256*/, BB_DS = 13 /*256*/,
  BB_PFREE = 14 /* dV(i) retired: P image free */, BB_DSFREE = 15 /* dK(i), dQ(i) retired: dS image free, iteration complete */,
  BB_DVDR = 16 /*256: dV accumulator drained*/, BB_DKDR = 17 /*256*/, BB_DQDR = 18 /*[2] 256*/, BB_COUNT = 20
};

__global__ void __launch_bounds__(kWsThreads, 1)
attn_bwd_ws(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQ,
            const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
            const __grid_constant__ WsParams a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
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
      for (int i = 0; i < n; ++i) {
        const int hl = i / per_head, rem = i - hl * per_head, kt = rem / QT, qt = rem - kt * QT;
        const int head = h_begin + hl, b = head / a.H, h = head - b * a.H;
        if (qt == 0) {                             // a new key tile: K, V into stage kc & 1
          const int kc = hl * KT + kt, st = kc & 1;
          if (kc >= 2) mbar_wait(bar(BB_KVFREE + st), static_cast<uint32_t>(((kc >> 1) - 1) & 1));
          mbar_arrive_expect_tx(bar(BB_KV + st), 2 * kChunkBytes);
          tma_load_3d(&tmK, bar(BB_KV + st), sKV_u + st * 2 * kChunkBytes, h * DH, kt * 128, b);
          tma_load_3d(&tmV, bar(BB_KV + st), sKV_u + st * 2 * kChunkBytes + kChunkBytes, h * DH, kt * 128, b);
        }
        if (kt == 0) {                             // first use of this query tile: Q, dO into stage qc & 1, O into its tile
          const int qc = hl * QT + qt, st = qc & 1;
          if (qc >= 2) mbar_wait(bar(BB_QFREE + st), static_cast<uint32_t>(((qc >> 1) - 1) & 1));
          mbar_arrive_expect_tx(bar(BB_QDO + st), 2 * kChunkBytes);
          tma_load_3d(&tmQ, bar(BB_QDO + st), sQD_u + st * 2 * kChunkBytes, h * DH, qt * 128, b);
          tma_load_3d(&tmDO, bar(BB_QDO + st), sQD_u + st * 2 * kChunkBytes + kChunkBytes, h * DH, qt * 128, b);
          if (qc >= 1) mbar_wait(bar(BB_OFREE), static_cast<uint32_t>((qc - 1) & 1));
          mbar_arrive_expect_tx(bar(BB_O), kChunkBytes);
          tma_load_3d(&tmO, bar(BB_O), sO_u, h * DH, qt * 128, b);
        }
      }
    }
  } else if (warp == 9) {
    // =============================== tcgen05 issuer ===============================
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, DH, true, true);     // dV, dK: A = P^T / dS^T (MN-major), B MN-major
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, DH, false, true);   // dQ: A = dS (K-major), B = K (MN-major)
      // operands of iteration i
      auto coords = [&](int i, int& hl, int& kt, int& qt) {
        hl = i / per_head;
        const int rem = i - hl * per_head;
        kt = rem / QT;
        qt = rem - kt * QT;
      };
      auto nkp_of = [&](int kt) { const int left = a.N - kt * 128; return left >= 128 ? 128 : ((left + 15) & ~15); };
      // S(i) = Q K^T and dP(i) = dO V^T  (N = this key tile's padded key count)
      auto wait_operands = [&](int i) {
        int hl, kt, qt;
        coords(i, hl, kt, qt);
        const int kc = hl * KT + kt, qc = hl * QT + qt;
        if (qt == 0) mbar_wait(bar(BB_KV + (kc & 1)), static_cast<uint32_t>((kc >> 1) & 1));
        if (kt == 0) mbar_wait(bar(BB_QDO + (qc & 1)), static_cast<uint32_t>((qc >> 1) & 1));
        tc_fence_after();
      };
      auto issue_s = [&](int i) {
        int hl, kt, qt;
        coords(i, hl, kt, qt);
        const int kc = hl * KT + kt, qc = hl * QT + qt;
        const uint32_t sK = sKV_u + (kc & 1) * 2 * kChunkBytes, sQ = sQD_u + (qc & 1) * 2 * kChunkBytes;
        const uint32_t idesc = umma_idesc_bf16(128, nkp_of(kt), false, false);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss(tmem_base + T_S, umma_smem_desc_sw128(sQ + k * 32, 16, 1024), umma_smem_desc_sw128(sK + k * 32, 16, 1024),
                       idesc, k > 0 ? 1u : 0u);
        umma_commit(bar(BB_S));
      };
      auto issue_dp = [&](int i) {
        int hl, kt, qt;
        coords(i, hl, kt, qt);
        const int kc = hl * KT + kt, qc = hl * QT + qt;
        const uint32_t sV = sKV_u + (kc & 1) * 2 * kChunkBytes + kChunkBytes, sDO = sQD_u + (qc & 1) * 2 * kChunkBytes + kChunkBytes;
        const uint32_t idesc = umma_idesc_bf16(128, nkp_of(kt), false, false);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss(tmem_base + T_DP, umma_smem_desc_sw128(sDO + k * 32, 16, 1024), umma_smem_desc_sw128(sV + k * 32, 16, 1024),
                       idesc, k > 0 ? 1u : 0u);
        umma_commit(bar(BB_DP));
      };
      wait_operands(0);
      issue_s(0);
      issue_dp(0);
      for (int i = 0; i < n; ++i) {
        int hl, kt, qt;
        coords(i, hl, kt, qt);
        const int kc = hl * KT + kt, qc = hl * QT + qt;
        const uint32_t ph = static_cast<uint32_t>(i & 1);
        const uint32_t sK = sKV_u + (kc & 1) * 2 * kChunkBytes;
        const uint32_t sQ = sQD_u + (qc & 1) * 2 * kChunkBytes, sDO = sQ + kChunkBytes;
        const int nkp = nkp_of(kt);
        // ---- P(i) is in shared memory (and S has been read): S(i+1), then dV += P^T dO
        mbar_wait(bar(BB_P), ph);
        tc_fence_after();
        if (i + 1 < n) { wait_operands(i + 1); issue_s(i + 1); }
        if (qt == 0 && kc > 0) { mbar_wait(bar(BB_DVDR), static_cast<uint32_t>((kc - 1) & 1)); tc_fence_after(); }
#pragma unroll
        for (int k = 0; k < 8; ++k)        // K = 128 query rows; A = P^T (MN-major image of sP), B = dO (MN-major)
          umma_bf16_ss(tmem_base + T_DV, umma_smem_desc_sw128(sP_u + k * 2048, kChunkBytes, 1024),
                       umma_smem_desc_sw128(sDO + k * 2048, 8192, 1024), idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar(BB_PFREE));
        // ---- dS(i) is in shared memory (and dP has been read): dP(i+1), then dK += dS^T Q, dQ += dS K
        mbar_wait(bar(BB_DS), ph);
        tc_fence_after();
        if (i + 1 < n) issue_dp(i + 1);
        if (qt == 0 && kc > 0) { mbar_wait(bar(BB_DKDR), static_cast<uint32_t>((kc - 1) & 1)); tc_fence_after(); }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tmem_base + T_DK, umma_smem_desc_sw128(sDS_u + k * 2048, kChunkBytes, 1024),
                       umma_smem_desc_sw128(sQ + k * 2048, 8192, 1024), idesc_t, (qt > 0 || k > 0) ? 1u : 0u);
        if (kt == 0 && hl > 0) { mbar_wait(bar(BB_DQDR + qt), static_cast<uint32_t>((hl - 1) & 1)); tc_fence_after(); }
        const int nks = nkp >> 4;
        for (int t16 = 0; t16 < nks; ++t16)   // K = this tile's keys; A = dS (K-major over keys), B = K (MN-major: keys x dh)
          umma_bf16_ss(tmem_base + T_DQ + static_cast<uint32_t>(qt * 64),
                       umma_smem_desc_sw128(sDS_u + (t16 >> 2) * kChunkBytes + (t16 & 3) * 32, 16, 1024),
                       umma_smem_desc_sw128(sK + t16 * 2048, 8192, 1024), idesc_dq, (kt > 0 || t16 > 0) ? 1u : 0u);
        umma_commit(bar(BB_DSFREE));
        if (qt == QT - 1) umma_commit(bar(BB_KVFREE + (kc & 1)));   // last user of this K, V stage
        if (kt == KT - 1) umma_commit(bar(BB_QFREE + (qc & 1)));    // last user of this Q, dO stage
      }
    }
  } else {
    // =============================== CUDA-core warps: P, dS, drains ===============================
    const int half = warp >> 2;                    // two threads per query row; `half` picks the 16-key units
    const int r = (warp & 3) * 32 + lane;          // query row within the tile == TMEM lane
    const bool elected = (tid == 0);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float lse2_q[2] = {0.f, 0.f}, d_q[2] = {0.f, 0.f};   // per query tile: LSE * log2(e) and D_i of this thread's row
    // accumulators waiting to be drained: (image, head) and tile they belong to
    bool pend_dv = false, pend_dk = false, pend_dq[2] = {false, false};
    int pend_kv_b = 0, pend_kv_h = 0, pend_kv_kt = 0, pend_dq_b[2] = {0, 0}, pend_dq_h[2] = {0, 0};

    // one accumulator tile: TMEM -> registers (the accumulator is released) -> bf16 staging tile -> TMA store
    auto drain = [&](uint32_t tcol, int drained_bar, float scale, const CUtensorMap* tm, int row0, int b, int h) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + tcol + static_cast<uint32_t>(half * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(drained_bar));
      if (elected) bulk_wait_read<0>();            // the previous store has finished reading the staging tile
      named_bar_sync(1, 256);
      stage_row32_bf16(sStage_u, r, half, v, scale);
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (elected) {
        tma_store_3d(tm, sStage_u, h * DH, row0, b);
        bulk_commit();
      }
    };
    auto drain_dv = [&]() {
      drain(T_DV, BB_DVDR, 1.0f, &tmDV, pend_kv_kt * 128, pend_kv_b, pend_kv_h);
      pend_dv = false;
    };
    auto drain_dk = [&]() {
      drain(T_DK, BB_DKDR, a.scale, &tmDK, pend_kv_kt * 128, pend_kv_b, pend_kv_h);
      pend_dk = false;
    };
    auto drain_dq = [&](int q) {
      drain(T_DQ + static_cast<uint32_t>(q * 64), BB_DQDR + q, a.scale, &tmDQ, q * 128, pend_dq_b[q], pend_dq_h[q]);
      pend_dq[q] = false;
    };

    for (int i = 0; i < n; ++i) {
      const int hl = i / per_head, rem = i - hl * per_head, kt = rem / QT, qt = rem - kt * QT;
      const int head = h_begin + hl, b = head / a.H, h = head - b * a.H;
      const int qc = hl * QT + qt;
      const uint32_t ph = static_cast<uint32_t>(i & 1);
      const int nk_valid = min(128, a.N - kt * 128);          // keys of this tile that exist
      const int nunits = (nk_valid + 15) >> 4;                // 16-key units carrying at least one key
      const int umid = (nunits + 1) >> 1;
      const int ub = half ? umid : 0, ue = half ? nunits : umid;   // this thread's units (<= 4)
      if (kt == 0) {
        // D_i = rowsum(dO * O) of this query tile from the swizzled tiles; LSE of this thread's row
        const int row = qt * 128 + r;
        const float lse = (row < a.N) ? a.lse[(static_cast<long long>(b) * a.H + h) * a.N + row] : INFINITY;
        mbar_wait(bar(BB_QDO + (qc & 1)), static_cast<uint32_t>((qc >> 1) & 1));
        mbar_wait(bar(BB_O), static_cast<uint32_t>(qc & 1));
        const uint32_t sDO = sQD_u + (qc & 1) * 2 * kChunkBytes + kChunkBytes;
        float part = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t off = swz_unit(r, half * 4 + u);
          part += dot8_bf16(ld_shared_v4(sO_u + off), ld_shared_v4(sDO + off));
        }
        red[half * 128 + r] = part;
        named_bar_sync(1, 256);
        const float di = red[r] + red[128 + r];
        mbar_arrive(bar(BB_OFREE));                // the O tile may be replaced
        named_bar_sync(1, 256);                    // red[] may be rewritten by the next query tile
        if (qt == 0) { d_q[0] = di; lse2_q[0] = lse * kLog2e; } else { d_q[1] = di; lse2_q[1] = lse * kLog2e; }
      }
      const float lse2 = qt == 0 ? lse2_q[0] : lse2_q[1];
      const float Di = qt == 0 ? d_q[0] : d_q[1];
      // ---- P = exp2(S*c - LSE*log2e) for this thread's units -> bf16 -> sP (rows >= N: LSE = +inf -> 0; keys >= N -> 0)
      mbar_wait(bar(BB_S), ph);
      if (i >= 1) mbar_wait(bar(BB_PFREE), static_cast<uint32_t>((i - 1) & 1));   // dV(i-1) has read the previous P
      tc_fence_after();
      uint32_t pk[4][8];
#pragma unroll
      for (int ul = 0; ul < 4; ++ul) {
        const int u = ub + ul;
        if (u < ue) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(trow + T_S + static_cast<uint32_t>(u * 16), v);
          tmem_ld_wait();
          const int c0 = u * 16;
          const bool interior = (c0 + 16 <= nk_valid);
          const uint64_t sc2 = pk2(a.scale_log2), nl2 = pk2(-lse2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x0, x1;
            upk2(fma2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc2, nl2), x0, x1);
            float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
            if (!interior) {
              e0 = (c0 + 2 * j < nk_valid) ? e0 : 0.f;
              e1 = (c0 + 2 * j + 1 < nk_valid) ? e1 : 0.f;
            }
            pk[ul][j] = pack_bf16x2(e0, e1);
          }
          const uint32_t base = sP_u + (c0 >> 6) * kChunkBytes;
          const int u8 = (c0 & 63) >> 3;
          st_shared_v4(base + swz_unit(r, u8), pk[ul][0], pk[ul][1], pk[ul][2], pk[ul][3]);
          st_shared_v4(base + swz_unit(r, u8 + 1), pk[ul][4], pk[ul][5], pk[ul][6], pk[ul][7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(BB_P));
      // ---- slot after P: the MMAs of iteration i-1 have retired by now; drain what they completed
      if (i >= 1) { mbar_wait(bar(BB_DSFREE), static_cast<uint32_t>((i - 1) & 1)); tc_fence_after(); }
      if (pend_dv) drain_dv();
      else if (pend_dq[qt ^ 1]) drain_dq(qt ^ 1);
      // ---- dS / c = P * (dP - D) -> bf16 -> sdS  (the softmax scale c is applied when dQ / dK are drained)
      mbar_wait(bar(BB_DP), ph);
      tc_fence_after();
#pragma unroll
      for (int ul = 0; ul < 4; ++ul) {
        const int u = ub + ul;
        if (u < ue) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(trow + T_DP + static_cast<uint32_t>(u * 16), v);
          tmem_ld_wait();
          uint32_t ds[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d0 = bf16_lo(pk[ul][j]) * (__uint_as_float(v[2 * j]) - Di);
            const float d1 = bf16_hi(pk[ul][j]) * (__uint_as_float(v[2 * j + 1]) - Di);
            ds[j] = pack_bf16x2(d0, d1);
          }
          const int c0 = u * 16;
          const uint32_t base = sDS_u + (c0 >> 6) * kChunkBytes;
          const int u8 = (c0 & 63) >> 3;
          st_shared_v4(base + swz_unit(r, u8), ds[0], ds[1], ds[2], ds[3]);
          st_shared_v4(base + swz_unit(r, u8 + 1), ds[4], ds[5], ds[6], ds[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(BB_DS));
      // ---- slot after dS
      if (pend_dk) drain_dk();
      if (pend_dq[qt]) drain_dq(qt);
      // what this iteration completes (drained one slot later, once its MMAs have retired)
      if (qt == QT - 1) { pend_dv = pend_dk = true; pend_kv_b = b; pend_kv_h = h; pend_kv_kt = kt; }
      if (kt == KT - 1) {
        if (qt == 0) { pend_dq[0] = true; pend_dq_b[0] = b; pend_dq_h[0] = h; }
        else { pend_dq[1] = true; pend_dq_b[1] = b; pend_dq_h[1] = h; }
      }
    }
    if (n > 0) {
      mbar_wait(bar(BB_DSFREE), static_cast<uint32_t>((n - 1) & 1));
      tc_fence_after();
      if (pend_dv) drain_dv();
      if (pend_dk) drain_dk();
      if (pend_dq[0]) drain_dq(0);
      if (pend_dq[1]) drain_dq(1);
    }
    if (elected) bulk_wait_all();                  // shared memory must outlive the reads of the last store
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

extern "C" int vitb_attn_ws_supported(int head_dim, int Nq, int Nk) {
  return head_dim == DH && Nq == Nk && Nk >= 1 && Nk <= 256;
}

extern "C" int vitb_attn_fwd_ws(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  st = check_ws(p, "attn_fwd_ws");
  if (st != VITB_OK) return st;
  if (p->B == 0) return VITB_OK;
  const int N = p->Nk, NK = (N + 15) & ~15;
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
  const int kv_bytes = NK * 128, pchunks = (NK + 63) / 64;
  // K | V | 2 query tiles | 2 P images | barriers + TMEM slot | alignment slack
  const int smem = 2 * kv_bytes + 2 * kChunkBytes + 2 * pchunks * kChunkBytes + 256 + 1024;
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_fwd_ws: %d B of shared memory", smem);
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int sms = vitb_num_sms();
  const int grid = a.total < sms ? a.total : sms;
  VITB_CUDA_CHECK(vitb_launch(attn_fwd_ws, dim3(grid), dim3(kWsThreads), smem, reinterpret_cast<cudaStream_t>(stream_), tq, tk,
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
  if ((st = head_map(&tdq, p->dq, p->H, N, p->B, p->dq_row_stride, p->dq_batch_stride, 128)) != VITB_OK) return st;
  if ((st = head_map(&tdk, p->dk, p->H, N, p->B, p->dk_row_stride, p->dk_batch_stride, 128)) != VITB_OK) return st;
  if ((st = head_map(&tdv, p->dv, p->H, N, p->B, p->dv_row_stride, p->dv_batch_stride, 128)) != VITB_OK) return st;
  WsParams a{};
  a.N = N; a.NK = (N + 15) & ~15; a.H = p->H;
  a.qtiles = (N + 127) / 128;
  a.ktiles = a.qtiles;
  a.total = p->B * p->H;
  a.scale = 1.0f / sqrtf((float)DH);
  a.scale_log2 = a.scale * kLog2e;
  a.lse = p->lse;
  // 14 tiles of 16 KiB | barriers + TMEM slot + D_i partials | alignment slack
  const int smem = 14 * kChunkBytes + 8 * (BB_COUNT + 2) + 2 * 128 * 4 + 1024;
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_ws: %d B of shared memory", smem);
  VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int sms = vitb_num_sms();
  const int grid = a.total < sms ? a.total : sms;
  VITB_CUDA_CHECK(vitb_launch(attn_bwd_ws, dim3(grid), dim3(kWsThreads), smem, reinterpret_cast<cudaStream_t>(stream_), tq, tk,
                              tv, tdo, to, tdq, tdk, tdv, a));
  VITB_LAUNCH_CHECK("attn_bwd_ws");
  return VITB_OK;
}
