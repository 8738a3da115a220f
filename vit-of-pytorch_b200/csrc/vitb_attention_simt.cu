// vitb_attention_simt.cu — CUDA-core (fp32 math) scaled-dot-product attention, forward and backward.
//
// Role on the path: the fp32 PARITY mode of SelfAttention (src/model.py:83-101) / Attention
// (res-vit/model.py:237-299), where probabilities must stay fp32 to meet the rel-1e-4 bar, and the
// head shapes the tcgen05 kernel (vitb_attention_tc.cu) does not take (head_dim != 64, or more
// than 256 keys).  Supports different query and key counts (res-vit's asymmetric eval attention,
// res-vit/model.py:503-516) through an optional per-image query row list.
//
//   S = (Q K^T) / sqrt(dh);  P = softmax(S);  O = P V;  LSE kept for the backward.
//
// Layout: element (b, n, h, d) of q/k/v/o lives at base + b*batch_stride + n*row_stride + h*dh + d.
// K and V of one (image, head) sit transposed in shared memory with an ODD row stride so that both
// "lanes over keys" and "lanes over head-dim" accesses are bank-conflict free.
#include "../../include/vitb200.h"
#include "vitb_common.cuh"

namespace {
using namespace vitb;

constexpr int kWarps = 8;
constexpr int kMaxKT = 10;  // up to 320 keys per row in registers (forward)

template <bool BF16>
__device__ __forceinline__ float ldf(const void* p, long long i) {
  if constexpr (BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  else return reinterpret_cast<const float*>(p)[i];
}
template <bool BF16>
__device__ __forceinline__ void stf(void* p, long long i, float v) {
  if constexpr (BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

struct AttnDev {
  const void *q, *k, *v;
  void* o;
  float* lse;  // [B, H, Nq]
  long long q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  int B, H, Nq, Nk, dh;
  float scale;  // 1/sqrt(dh)
  // backward
  const void* dout;
  long long do_bs, do_rs;
  float *dq, *dk, *dv;  // fp32 accumulators laid out like q/k/v with their own strides
  long long dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  int key_chunk;
};

// grid: (B*H, q_splits).  smem: Kt[dh][Ns], Vt[dh][Ns], per-warp q[dh] and p[Ns].
template <bool BF16>
__global__ void __launch_bounds__(kWarps * 32)
attn_fwd_simt(const AttnDev a) {
  extern __shared__ float sm[];
  const int Ns = a.Nk | 1;  // odd stride
  float* Kt = sm;
  float* Vt = Kt + a.dh * Ns;
  float* wq = Vt + a.dh * Ns;          // [kWarps][dh]
  float* wp = wq + kWarps * a.dh;      // [kWarps][Ns]
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < a.Nk * a.dh; i += blockDim.x) {
    const int j = i / a.dh, d = i - j * a.dh;
    Kt[d * Ns + j] = ldf<BF16>(a.k, b * a.k_bs + j * a.k_rs + h * a.dh + d);
    Vt[d * Ns + j] = ldf<BF16>(a.v, b * a.v_bs + j * a.v_rs + h * a.dh + d);
  }
  __syncthreads();
  float* q = wq + warp * a.dh;
  float* p = wp + warp * Ns;
  const int rows_per = (a.Nq + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(a.Nq, r0 + rows_per);
  for (int i = r0 + warp; i < r1; i += kWarps) {
    for (int d = lane; d < a.dh; d += 32) q[d] = ldf<BF16>(a.q, b * a.q_bs + i * a.q_rs + h * a.dh + d);
    __syncwarp();
    float s[kMaxKT];
#pragma unroll
    for (int t = 0; t < kMaxKT; ++t) s[t] = 0.f;
    for (int d = 0; d < a.dh; ++d) {
      const float qd = q[d];
      const float* kr = Kt + d * Ns;
#pragma unroll
      for (int t = 0; t < kMaxKT; ++t) {
        const int j = lane + 32 * t;
        if (j < a.Nk) s[t] = fmaf(qd, kr[j], s[t]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < kMaxKT; ++t) {
      s[t] *= a.scale;
      if (lane + 32 * t < a.Nk) mx = fmaxf(mx, s[t]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxKT; ++t) {
      const int j = lane + 32 * t;
      if (j < a.Nk) { s[t] = expf(s[t] - mx); sum += s[t]; }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < kMaxKT; ++t) {
      const int j = lane + 32 * t;
      if (j < a.Nk) p[j] = s[t] * inv;
    }
    if (lane == 0 && a.lse) a.lse[(static_cast<long long>(b) * a.H + h) * a.Nq + i] = mx + logf(sum);
    __syncwarp();
    for (int d = lane; d < a.dh; d += 32) {
      const float* vr = Vt + d * Ns;
      float acc = 0.f;
      for (int j = 0; j < a.Nk; ++j) acc = fmaf(p[j], vr[j], acc);
      stf<BF16>(a.o, b * a.o_bs + i * a.o_rs + h * a.dh + d, acc);
    }
    __syncwarp();
  }
}

// grid: (B*H, key_chunks).  Each block owns keys [c0,c1): dK/dV for them are complete in-block
// (accumulated in shared memory), dQ contributions go to global fp32 atomics (dq zeroed by caller).
template <bool BF16>
__global__ void __launch_bounds__(kWarps * 32)
attn_bwd_simt(const AttnDev a) {
  extern __shared__ float sm[];
  const int c0 = blockIdx.y * a.key_chunk;
  const int nk = min(a.key_chunk, a.Nk - c0);
  const int Ns = a.key_chunk | 1;
  float* Kt = sm;
  float* Vt = Kt + a.dh * Ns;
  float* dKt = Vt + a.dh * Ns;
  float* dVt = dKt + a.dh * Ns;
  float* wrow = dVt + a.dh * Ns;  // per warp: q[dh], do[dh], p[Ns], ds[Ns]
  const int per_warp = 2 * a.dh + 2 * Ns;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < a.dh * Ns; i += blockDim.x) { dKt[i] = 0.f; dVt[i] = 0.f; }
  for (int i = threadIdx.x; i < nk * a.dh; i += blockDim.x) {
    const int j = i / a.dh, d = i - j * a.dh;
    Kt[d * Ns + j] = ldf<BF16>(a.k, b * a.k_bs + (c0 + j) * a.k_rs + h * a.dh + d);
    Vt[d * Ns + j] = ldf<BF16>(a.v, b * a.v_bs + (c0 + j) * a.v_rs + h * a.dh + d);
  }
  __syncthreads();
  float* q = wrow + warp * per_warp;
  float* dO = q + a.dh;
  float* p = dO + a.dh;
  float* ds = p + Ns;
  for (int i = warp; i < a.Nq; i += kWarps) {
    float dsum = 0.f;
    for (int d = lane; d < a.dh; d += 32) {
      const float qv = ldf<BF16>(a.q, b * a.q_bs + i * a.q_rs + h * a.dh + d);
      const float gv = ldf<BF16>(a.dout, b * a.do_bs + i * a.do_rs + h * a.dh + d);
      const float ov = ldf<BF16>(a.o, b * a.o_bs + i * a.o_rs + h * a.dh + d);
      q[d] = qv;
      dO[d] = gv;
      dsum = fmaf(gv, ov, dsum);
    }
    const float Di = warp_sum(dsum);
    const float lse = a.lse[(static_cast<long long>(b) * a.H + h) * a.Nq + i];
    __syncwarp();
    for (int j = lane; j < nk; j += 32) {
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < a.dh; ++d) {
        s = fmaf(q[d], Kt[d * Ns + j], s);
        dp = fmaf(dO[d], Vt[d * Ns + j], dp);
      }
      const float pj = expf(s * a.scale - lse);
      p[j] = pj;
      ds[j] = pj * (dp - Di) * a.scale;
    }
    __syncwarp();
    for (int d = lane; d < a.dh; d += 32) {
      const float* kr = Kt + d * Ns;
      float* dkr = dKt + d * Ns;
      float* dvr = dVt + d * Ns;
      const float qd = q[d], gd = dO[d];
      float acc = 0.f;
      for (int j = 0; j < nk; ++j) {
        const float dsj = ds[j];
        acc = fmaf(dsj, kr[j], acc);
        atomicAdd(dkr + j, dsj * qd);
        atomicAdd(dvr + j, p[j] * gd);
      }
      atomicAdd(a.dq + b * a.dq_bs + i * a.dq_rs + h * a.dh + d, acc);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nk * a.dh; i += blockDim.x) {
    const int j = i / a.dh, d = i - j * a.dh;
    a.dk[b * a.dk_bs + (c0 + j) * a.dk_rs + h * a.dh + d] = dKt[d * Ns + j];
    a.dv[b * a.dv_bs + (c0 + j) * a.dv_rs + h * a.dh + d] = dVt[d * Ns + j];
  }
}


// ================================================================================================
// One query per (image, head): the class-token-only attention of the last encoder block
// (EncoderBlock.forward_row0; the reference classifies feat[:, 0] only, src/model.py:155,210).
// The general kernels above stage K and V transposed in shared memory and leave one warp working when
// Nq = 1 (round-1 launch list: 0.22 ms forward + 0.53 ms backward per step for a few MFLOP).  Here every
// K / V row is read exactly once, straight from HBM, a warp per key with lanes over the head dimension
// (4- or 8-byte loads, 128 B or more per warp), so both kernels are bound by the K/V read and the dK/dV write.
// ================================================================================================
constexpr int kQ1Warps = 4;

template <bool BF16>
__device__ __forceinline__ float2 ld2(const void* p, long long i) {   // elements i, i+1 (i even)
  if constexpr (BF16) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    return make_float2(bf16_lo(u), bf16_hi(u));
  } else {
    return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p) + i);
  }
}

__device__ __forceinline__ float block_reduce_q1(float v, float* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();                 // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kQ1Warps; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// grid: B*H blocks of 4 warps.  smem: q[dh] | s[Nk] | part[4][dh] | red[4]
template <bool BF16>
__global__ void __launch_bounds__(kQ1Warps * 32)
attn_q1_fwd(const AttnDev a) {
  extern __shared__ float sm[];
  float* qs = sm;
  float* sc = qs + a.dh;
  float* part = sc + a.Nk;
  float* red = part + kQ1Warps * a.dh;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d = threadIdx.x; d < a.dh; d += blockDim.x) qs[d] = ldf<BF16>(a.q, b * a.q_bs + h * a.dh + d);
  __syncthreads();
  const long long kb = b * a.k_bs + h * a.dh, vb = b * a.v_bs + h * a.dh;
  for (int j = warp; j < a.Nk; j += kQ1Warps) {
    float acc = 0.f;
    for (int d = 2 * lane; d < a.dh; d += 64) {
      const float2 kv = ld2<BF16>(a.k, kb + j * a.k_rs + d);
      acc = fmaf(qs[d], kv.x, fmaf(qs[d + 1], kv.y, acc));
    }
    acc = warp_sum(acc);
    if (lane == 0) sc[j] = acc * a.scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < a.Nk; j += blockDim.x) mx = fmaxf(mx, sc[j]);
  mx = block_reduce_q1(mx, red, true);
  float sum = 0.f;
  for (int j = threadIdx.x; j < a.Nk; j += blockDim.x) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = block_reduce_q1(sum, red, false);   // its barriers also publish the exponentials
  const float inv = 1.0f / sum;
  if (threadIdx.x == 0 && a.lse) a.lse[static_cast<long long>(b) * a.H + h] = mx + logf(sum);
  // o = sum_j p_j V[j]: warp w takes keys w, w+4, ...; lanes hold pairs of head-dim columns
  float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);   // columns 2*lane(+1) and 64 + 2*lane(+1)
  for (int j = warp; j < a.Nk; j += kQ1Warps) {
    const float pj = sc[j] * inv;
    const int d = 2 * lane;
    if (d < a.dh) { const float2 v = ld2<BF16>(a.v, vb + j * a.v_rs + d); o0.x = fmaf(pj, v.x, o0.x); o0.y = fmaf(pj, v.y, o0.y); }
    if (d + 64 < a.dh) { const float2 v = ld2<BF16>(a.v, vb + j * a.v_rs + d + 64); o1.x = fmaf(pj, v.x, o1.x); o1.y = fmaf(pj, v.y, o1.y); }
  }
  {
    const int d = 2 * lane;
    if (d < a.dh) { part[warp * a.dh + d] = o0.x; part[warp * a.dh + d + 1] = o0.y; }
    if (d + 64 < a.dh) { part[warp * a.dh + d + 64] = o1.x; part[warp * a.dh + d + 65] = o1.y; }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < a.dh; d += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kQ1Warps; ++w) acc += part[w * a.dh + d];
    stf<BF16>(a.o, b * a.o_bs + h * a.dh + d, acc);
  }
}

// grid: B*H blocks of 4 warps.  smem: q[dh] | do[dh] | part[4][dh] | red[4].  dq / dk / dv are fp32 (SIMT ABI).
template <bool BF16>
__global__ void __launch_bounds__(kQ1Warps * 32)
attn_q1_bwd(const AttnDev a) {
  extern __shared__ float sm[];
  float* qs = sm;
  float* gs = qs + a.dh;
  float* part = gs + a.dh;
  float* red = part + kQ1Warps * a.dh;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dsum = 0.f;
  for (int d = threadIdx.x; d < a.dh; d += blockDim.x) {
    const float qv = ldf<BF16>(a.q, b * a.q_bs + h * a.dh + d);
    const float gv = ldf<BF16>(a.dout, b * a.do_bs + h * a.dh + d);
    const float ov = ldf<BF16>(a.o, b * a.o_bs + h * a.dh + d);
    qs[d] = qv;
    gs[d] = gv;
    dsum = fmaf(gv, ov, dsum);
  }
  const float Di = block_reduce_q1(dsum, red, false);   // also publishes qs / gs
  const float lse = a.lse[static_cast<long long>(b) * a.H + h];
  const long long kb = b * a.k_bs + h * a.dh, vb = b * a.v_bs + h * a.dh;
  const long long dkb = b * a.dk_bs + h * a.dh, dvb = b * a.dv_bs + h * a.dh;
  const int d0 = 2 * lane, d1 = 2 * lane + 64;
  const bool has0 = d0 < a.dh, has1 = d1 < a.dh;
  float2 q0 = make_float2(0.f, 0.f), q1 = q0, g0 = q0, g1 = q0;
  if (has0) { q0 = make_float2(qs[d0], qs[d0 + 1]); g0 = make_float2(gs[d0], gs[d0 + 1]); }
  if (has1) { q1 = make_float2(qs[d1], qs[d1 + 1]); g1 = make_float2(gs[d1], gs[d1 + 1]); }
  float2 dq0 = make_float2(0.f, 0.f), dq1 = dq0;
  for (int j = warp; j < a.Nk; j += kQ1Warps) {
    float2 k0 = make_float2(0.f, 0.f), k1 = k0, v0 = k0, v1 = k0;
    if (has0) { k0 = ld2<BF16>(a.k, kb + j * a.k_rs + d0); v0 = ld2<BF16>(a.v, vb + j * a.v_rs + d0); }
    if (has1) { k1 = ld2<BF16>(a.k, kb + j * a.k_rs + d1); v1 = ld2<BF16>(a.v, vb + j * a.v_rs + d1); }
    float s = q0.x * k0.x + q0.y * k0.y + q1.x * k1.x + q1.y * k1.y;
    float dp = g0.x * v0.x + g0.y * v0.y + g1.x * v1.x + g1.y * v1.y;
    s = warp_sum(s);
    dp = warp_sum(dp);
    const float pj = expf(s * a.scale - lse);
    const float ds = pj * (dp - Di) * a.scale;
    if (has0) {
      *reinterpret_cast<float2*>(a.dk + dkb + j * a.dk_rs + d0) = make_float2(ds * q0.x, ds * q0.y);
      *reinterpret_cast<float2*>(a.dv + dvb + j * a.dv_rs + d0) = make_float2(pj * g0.x, pj * g0.y);
      dq0.x = fmaf(ds, k0.x, dq0.x); dq0.y = fmaf(ds, k0.y, dq0.y);
    }
    if (has1) {
      *reinterpret_cast<float2*>(a.dk + dkb + j * a.dk_rs + d1) = make_float2(ds * q1.x, ds * q1.y);
      *reinterpret_cast<float2*>(a.dv + dvb + j * a.dv_rs + d1) = make_float2(pj * g1.x, pj * g1.y);
      dq1.x = fmaf(ds, k1.x, dq1.x); dq1.y = fmaf(ds, k1.y, dq1.y);
    }
  }
  if (has0) { part[warp * a.dh + d0] = dq0.x; part[warp * a.dh + d0 + 1] = dq0.y; }
  if (has1) { part[warp * a.dh + d1] = dq1.x; part[warp * a.dh + d1 + 1] = dq1.y; }
  __syncthreads();
  for (int d = threadIdx.x; d < a.dh; d += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kQ1Warps; ++w) acc += part[w * a.dh + d];
    a.dq[b * a.dq_bs + h * a.dh + d] = acc;
  }
}

// ---- single-query attention, bf16, head_dim 64, <= 256 keys: the bandwidth-bound version --------------------------
// The two kernels above walk the keys one per warp and iteration (a 128-byte load, a warp reduction, the next key):
// latency-bound at 0.85 TB/s for the class-token-only last block of ViT-B/16 (profiles/launches_r02k_summary.txt).
// Here eight lanes own one key row (16 bytes = 8 head-dim columns each), a warp covers four keys per load instruction,
// and ALL of a thread's K and V loads (<= 16 + 16 of 16 bytes) are issued before the first one is consumed.
// grid: B*H blocks of 4 warps; key j is owned by slot j % 16 = warp * 4 + lane / 8, iteration j / 16.
constexpr int kQ1FastIt = 16;        // 16 iterations x 16 key slots = 256 keys

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ float dot8(const uint4& u, const float (&q)[8]) {
  float f[8];
  unpack8(u, f);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(f[i], q[i], s);
  return s;
}
__device__ __forceinline__ float sum8lanes(float v) {      // over the eight lanes that share a key row
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// smem: sc[256] | part[4][64] | red[4]
__global__ void __launch_bounds__(kQ1Warps * 32)
attn_q1_fwd64(const AttnDev a) {
  __shared__ float sc[kQ1FastIt * 16];
  __shared__ float part[kQ1Warps * 64];
  __shared__ float red[kQ1Warps];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp * 4 + (lane >> 3), c8 = (lane & 7) * 8;
  const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(a.k) + b * a.k_bs + h * 64 + c8;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(a.v) + b * a.v_bs + h * 64 + c8;
  uint4 kr[kQ1FastIt], vr[kQ1FastIt];
#pragma unroll
  for (int it = 0; it < kQ1FastIt; ++it) {
    const int j = it * 16 + slot;
    kr[it] = make_uint4(0u, 0u, 0u, 0u);
    if (j < a.Nk) kr[it] = *reinterpret_cast<const uint4*>(kp + j * a.k_rs);
  }
#pragma unroll
  for (int it = 0; it < kQ1FastIt; ++it) {
    const int j = it * 16 + slot;
    vr[it] = make_uint4(0u, 0u, 0u, 0u);
    if (j < a.Nk) vr[it] = *reinterpret_cast<const uint4*>(vp + j * a.v_rs);
  }
  float q[8];
  unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.q) + b * a.q_bs + h * 64 + c8), q);
  float mx = -INFINITY;
#pragma unroll
  for (int it = 0; it < kQ1FastIt; ++it) {
    const int j = it * 16 + slot;
    const float s = sum8lanes(dot8(kr[it], q)) * a.scale;
    if (j < a.Nk) {
      mx = fmaxf(mx, s);
      if ((lane & 7) == 0) sc[j] = s;
    }
  }
  mx = block_reduce_q1(mx, red, true);     // its barriers publish sc
  float sum = 0.f;
  for (int j = threadIdx.x; j < a.Nk; j += blockDim.x) sum += expf(sc[j] - mx);
  sum = block_reduce_q1(sum, red, false);
  const float inv = 1.0f / sum;
  if (threadIdx.x == 0 && a.lse) a.lse[static_cast<long long>(b) * a.H + h] = mx + logf(sum);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
  for (int it = 0; it < kQ1FastIt; ++it) {
    const int j = it * 16 + slot;
    if (j < a.Nk) {
      const float pj = expf(sc[j] - mx) * inv;
      float f[8];
      unpack8(vr[it], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(pj, f[i], o[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {              // the warp's four key slots
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) part[warp * 64 + c8 + i] = o[i];
  }
  __syncthreads();
  if (threadIdx.x < 32) {                    // two columns per lane
    const int d = 2 * threadIdx.x;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < kQ1Warps; ++w) { o0 += part[w * 64 + d]; o1 += part[w * 64 + d + 1]; }
    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.o) + b * a.o_bs + h * 64 + d) = pack_bf16x2(o0, o1);
  }
}

// dq / dk / dv are BF16 here (the tensor dtype), laid out through their own strides; dk and dv rows are written whole.
__global__ void __launch_bounds__(kQ1Warps * 32)
attn_q1_bwd64(const AttnDev a) {
  __shared__ float part[kQ1Warps * 64];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp * 4 + (lane >> 3), c8 = (lane & 7) * 8;
  const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(a.k) + b * a.k_bs + h * 64 + c8;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(a.v) + b * a.v_bs + h * 64 + c8;
  uint4 kr[kQ1FastIt], vr[kQ1FastIt];
#pragma unroll
  for (int it = 0; it < kQ1FastIt; ++it) {
    const int j = it * 16 + slot;
    kr[it] = make_uint4(0u, 0u, 0u, 0u);
    vr[it] = make_uint4(0u, 0u, 0u, 0u);
    if (j < a.Nk) {
      kr[it] = *reinterpret_cast<const uint4*>(kp + j * a.k_rs);
      vr[it] = *reinterpret_cast<const uint4*>(vp + j * a.v_rs);
    }
  }
  float q[8], g[8], ov[8];
  unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.q) + b * a.q_bs + h * 64 + c8), q);
  unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.dout) + b * a.do_bs + h * 64 + c8), g);
  unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.o) + b * a.o_bs + h * 64 + c8), ov);
  float di = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) di = fmaf(g[i], ov[i], di);
  const float Di = sum8lanes(di);           // every group of eight lanes holds the whole row
  const float lse = a.lse[static_cast<long long>(b) * a.H + h];
  __nv_bfloat16* dkp = reinterpret_cast<__nv_bfloat16*>(a.dk) + b * a.dk_bs + h * 64 + c8;
  __nv_bfloat16* dvp = reinterpret_cast<__nv_bfloat16*>(a.dv) + b * a.dv_bs + h * 64 + c8;
  float dq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i] = 0.f;
#pragma unroll
  for (int it = 0; it < kQ1FastIt; ++it) {
    const int j = it * 16 + slot;
    const float s = sum8lanes(dot8(kr[it], q));
    const float dp = sum8lanes(dot8(vr[it], g));
    if (j < a.Nk) {
      const float pj = expf(s * a.scale - lse);
      const float ds = pj * (dp - Di) * a.scale;
      float kf[8];
      unpack8(kr[it], kf);
      uint4 wk, wv;
      wk.x = pack_bf16x2(ds * q[0], ds * q[1]); wk.y = pack_bf16x2(ds * q[2], ds * q[3]);
      wk.z = pack_bf16x2(ds * q[4], ds * q[5]); wk.w = pack_bf16x2(ds * q[6], ds * q[7]);
      wv.x = pack_bf16x2(pj * g[0], pj * g[1]); wv.y = pack_bf16x2(pj * g[2], pj * g[3]);
      wv.z = pack_bf16x2(pj * g[4], pj * g[5]); wv.w = pack_bf16x2(pj * g[6], pj * g[7]);
      *reinterpret_cast<uint4*>(dkp + j * a.dk_rs) = wk;
      *reinterpret_cast<uint4*>(dvp + j * a.dv_rs) = wv;
#pragma unroll
      for (int i = 0; i < 8; ++i) dq[i] = fmaf(ds, kf[i], dq[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dq[i] += __shfl_xor_sync(0xffffffffu, dq[i], 8);
    dq[i] += __shfl_xor_sync(0xffffffffu, dq[i], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) part[warp * 64 + c8 + i] = dq[i];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int d = 2 * threadIdx.x;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < kQ1Warps; ++w) { o0 += part[w * 64 + d]; o1 += part[w * 64 + d + 1]; }
    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(a.dq) + b * a.dq_bs + h * 64 + d) = pack_bf16x2(o0, o1);
  }
}

// shapes of the two kernels above: one bf16 query per image, head_dim 64, <= 256 keys, 16-byte aligned rows
bool q1_fast_ok(const vitb_attn_params* p, bool grads) {
  auto al8 = [](long long v) { return v % 8 == 0; };
  auto p16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  bool ok = p->Nq == 1 && p->dtype == VITB_BF16 && p->head_dim == 64 && p->Nk >= 1 && p->Nk <= 16 * kQ1FastIt &&
            al8(p->q_batch_stride) && al8(p->k_batch_stride) && al8(p->k_row_stride) && al8(p->v_batch_stride) &&
            al8(p->v_row_stride) && al8(p->o_batch_stride) && p16(p->q) && p16(p->k) && p16(p->v) && p16(p->o);
  if (ok && grads)
    ok = al8(p->do_batch_stride) && al8(p->dq_batch_stride) && al8(p->dk_batch_stride) && al8(p->dk_row_stride) &&
         al8(p->dv_batch_stride) && al8(p->dv_row_stride) && p16(p->dout) && p16(p->dq) && p16(p->dk) && p16(p->dv);
  return ok;
}

// single-query shapes these kernels take: head_dim even and <= 128, 8-byte aligned fp32 gradient rows
bool q1_ok(const vitb_attn_params* p) {
  return p->Nq == 1 && p->head_dim % 2 == 0 && p->head_dim <= 128 && p->k_row_stride % 2 == 0 && p->v_row_stride % 2 == 0 &&
         p->k_batch_stride % 2 == 0 && p->v_batch_stride % 2 == 0;
}

}  // namespace

extern "C" int vitb_attn_fwd_simt(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(p && p->struct_bytes == (int)sizeof(vitb_attn_params), VITB_ERR_BAD_ARG, "attn_fwd_simt: ABI mismatch");
  if (p->B == 0 || p->Nq == 0) return VITB_OK;
  VITB_REQUIRE(p->q && p->k && p->v && p->o, VITB_ERR_BAD_ARG, "attn_fwd_simt: null tensor");
  VITB_REQUIRE(p->Nk >= 1, VITB_ERR_UNSUPPORTED_SHAPE, "attn_fwd_simt: Nk=%d", p->Nk);
  AttnDev a{};
  a.q = p->q; a.k = p->k; a.v = p->v; a.o = p->o; a.lse = p->lse;
  a.q_bs = p->q_batch_stride; a.q_rs = p->q_row_stride; a.k_bs = p->k_batch_stride; a.k_rs = p->k_row_stride;
  a.v_bs = p->v_batch_stride; a.v_rs = p->v_row_stride; a.o_bs = p->o_batch_stride; a.o_rs = p->o_row_stride;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.dh = p->head_dim;
  a.scale = 1.0f / sqrtf((float)p->head_dim);
  if (q1_fast_ok(p, false)) {
    attn_q1_fwd64<<<p->B * p->H, kQ1Warps * 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(a);
    VITB_LAUNCH_CHECK("attn_q1_fwd64");
    return VITB_OK;
  }
  if (q1_ok(p)) {
    const size_t sm1 = sizeof(float) * ((size_t)a.dh + a.Nk + kQ1Warps * a.dh + kQ1Warps);
    VITB_REQUIRE(sm1 <= 200 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_q1_fwd: Nk=%d", p->Nk);
    cudaStream_t s1 = reinterpret_cast<cudaStream_t>(stream_);
    if (p->dtype == VITB_BF16) {
      VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_q1_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
      attn_q1_fwd<true><<<p->B * p->H, kQ1Warps * 32, sm1, s1>>>(a);
    } else {
      VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_q1_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
      attn_q1_fwd<false><<<p->B * p->H, kQ1Warps * 32, sm1, s1>>>(a);
    }
    VITB_LAUNCH_CHECK("attn_q1_fwd");
    return VITB_OK;
  }
  VITB_REQUIRE(p->Nk <= 32 * kMaxKT, VITB_ERR_UNSUPPORTED_SHAPE, "attn_fwd_simt: Nk=%d (max %d)", p->Nk, 32 * kMaxKT);
  const int Ns = p->Nk | 1;
  const size_t smem = sizeof(float) * (2 * (size_t)a.dh * Ns + kWarps * (a.dh + Ns));
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_fwd_simt: needs %zu B of shared memory", smem);
  int qsplit = (4 * vitb_num_sms() + p->B * p->H - 1) / (p->B * p->H);
  if (qsplit > (p->Nq + kWarps - 1) / kWarps) qsplit = (p->Nq + kWarps - 1) / kWarps;
  if (qsplit < 1) qsplit = 1;
  dim3 grid(p->B * p->H, qsplit);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (p->dtype == VITB_BF16) {
    VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_simt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt<true><<<grid, kWarps * 32, smem, stream>>>(a);
  } else {
    VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_simt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_simt<false><<<grid, kWarps * 32, smem, stream>>>(a);
  }
  VITB_LAUNCH_CHECK("attn_fwd_simt");
  return VITB_OK;
}

extern "C" int vitb_attn_bwd_simt(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(p && p->struct_bytes == (int)sizeof(vitb_attn_params), VITB_ERR_BAD_ARG, "attn_bwd_simt: ABI mismatch");
  if (p->B == 0 || p->Nq == 0) return VITB_OK;
  VITB_REQUIRE(p->q && p->k && p->v && p->o && p->lse && p->dout && p->dq && p->dk && p->dv, VITB_ERR_BAD_ARG,
               "attn_bwd_simt: null tensor");
  AttnDev a{};
  a.q = p->q; a.k = p->k; a.v = p->v; a.o = p->o; a.lse = p->lse;
  a.q_bs = p->q_batch_stride; a.q_rs = p->q_row_stride; a.k_bs = p->k_batch_stride; a.k_rs = p->k_row_stride;
  a.v_bs = p->v_batch_stride; a.v_rs = p->v_row_stride; a.o_bs = p->o_batch_stride; a.o_rs = p->o_row_stride;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.dh = p->head_dim;
  a.scale = 1.0f / sqrtf((float)p->head_dim);
  a.dout = p->dout; a.do_bs = p->do_batch_stride; a.do_rs = p->do_row_stride;
  a.dq = reinterpret_cast<float*>(p->dq); a.dk = reinterpret_cast<float*>(p->dk); a.dv = reinterpret_cast<float*>(p->dv);
  a.dq_bs = p->dq_batch_stride; a.dq_rs = p->dq_row_stride; a.dk_bs = p->dk_batch_stride; a.dk_rs = p->dk_row_stride;
  a.dv_bs = p->dv_batch_stride; a.dv_rs = p->dv_row_stride;
  if (q1_ok(p) && p->dk_row_stride % 2 == 0 && p->dv_row_stride % 2 == 0 && p->dk_batch_stride % 2 == 0 &&
      p->dv_batch_stride % 2 == 0) {
    const size_t sm1 = sizeof(float) * (2 * (size_t)a.dh + kQ1Warps * a.dh + kQ1Warps);
    cudaStream_t s1 = reinterpret_cast<cudaStream_t>(stream_);
    if (p->dtype == VITB_BF16) attn_q1_bwd<true><<<p->B * p->H, kQ1Warps * 32, sm1, s1>>>(a);
    else attn_q1_bwd<false><<<p->B * p->H, kQ1Warps * 32, sm1, s1>>>(a);
    VITB_LAUNCH_CHECK("attn_q1_bwd");
    return VITB_OK;
  }
  // key chunk so that 4 transposed [dh][chunk] arrays + per-warp rows fit in shared memory
  int chunk = p->Nk;
  for (;;) {
    const int Ns = chunk | 1;
    const size_t smem = sizeof(float) * (4 * (size_t)a.dh * Ns + kWarps * (2 * a.dh + 2 * Ns));
    if (smem <= 200 * 1024 || chunk <= 32) break;
    chunk = (chunk + 1) / 2;
  }
  a.key_chunk = chunk;
  const int Ns = chunk | 1;
  const size_t smem = sizeof(float) * (4 * (size_t)a.dh * Ns + kWarps * (2 * a.dh + 2 * Ns));
  VITB_REQUIRE(smem <= 227 * 1024, VITB_ERR_UNSUPPORTED_SHAPE, "attn_bwd_simt: needs %zu B of shared memory", smem);
  dim3 grid(p->B * p->H, (p->Nk + chunk - 1) / chunk);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (p->dtype == VITB_BF16) {
    VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_simt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_simt<true><<<grid, kWarps * 32, smem, stream>>>(a);
  } else {
    VITB_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_simt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_simt<false><<<grid, kWarps * 32, smem, stream>>>(a);
  }
  VITB_LAUNCH_CHECK("attn_bwd_simt");
  return VITB_OK;
}

// Backward of single-query attention with gradients in the tensor dtype (bf16): the class-token-only last encoder block
// (EncoderBlock.forward_row0).  bf16, head_dim 64, Nq == 1, Nk <= 256, 16-byte aligned rows; dk / dv rows are written
// whole (no zeroing needed), dq is [B, 1, H*64].
extern "C" int vitb_attn_q1_supported(int head_dim, int Nk) { return head_dim == 64 && Nk >= 1 && Nk <= 16 * kQ1FastIt; }

extern "C" int vitb_attn_q1_bwd(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(p && p->struct_bytes == (int)sizeof(vitb_attn_params), VITB_ERR_BAD_ARG, "attn_q1_bwd: ABI mismatch");
  if (p->B == 0) return VITB_OK;
  VITB_REQUIRE(p->q && p->k && p->v && p->o && p->lse && p->dout && p->dq && p->dk && p->dv, VITB_ERR_BAD_ARG,
               "attn_q1_bwd: null tensor");
  VITB_REQUIRE(q1_fast_ok(p, true), VITB_ERR_UNSUPPORTED_SHAPE,
               "attn_q1_bwd: needs bf16, Nq 1, head_dim 64, <= %d keys, 16-byte aligned rows (Nq=%d dh=%d Nk=%d)",
               16 * kQ1FastIt, p->Nq, p->head_dim, p->Nk);
  AttnDev a{};
  a.q = p->q; a.k = p->k; a.v = p->v; a.o = p->o; a.lse = p->lse;
  a.q_bs = p->q_batch_stride; a.q_rs = p->q_row_stride; a.k_bs = p->k_batch_stride; a.k_rs = p->k_row_stride;
  a.v_bs = p->v_batch_stride; a.v_rs = p->v_row_stride; a.o_bs = p->o_batch_stride; a.o_rs = p->o_row_stride;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.dh = p->head_dim;
  a.scale = 1.0f / sqrtf((float)p->head_dim);
  a.dout = p->dout; a.do_bs = p->do_batch_stride; a.do_rs = p->do_row_stride;
  a.dq = reinterpret_cast<float*>(p->dq); a.dk = reinterpret_cast<float*>(p->dk); a.dv = reinterpret_cast<float*>(p->dv);
  a.dq_bs = p->dq_batch_stride; a.dq_rs = p->dq_row_stride; a.dk_bs = p->dk_batch_stride; a.dk_rs = p->dk_row_stride;
  a.dv_bs = p->dv_batch_stride; a.dv_rs = p->dv_row_stride;
  attn_q1_bwd64<<<p->B * p->H, kQ1Warps * 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(a);
  VITB_LAUNCH_CHECK("attn_q1_bwd64");
  return VITB_OK;
}
