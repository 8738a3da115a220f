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

}  // namespace

extern "C" int vitb_attn_fwd_simt(const vitb_attn_params* p, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(p && p->struct_bytes == (int)sizeof(vitb_attn_params), VITB_ERR_BAD_ARG, "attn_fwd_simt: ABI mismatch");
  if (p->B == 0 || p->Nq == 0) return VITB_OK;
  VITB_REQUIRE(p->q && p->k && p->v && p->o, VITB_ERR_BAD_ARG, "attn_fwd_simt: null tensor");
  VITB_REQUIRE(p->Nk >= 1 && p->Nk <= 32 * kMaxKT, VITB_ERR_UNSUPPORTED_SHAPE, "attn_fwd_simt: Nk=%d (max %d)", p->Nk, 32 * kMaxKT);
  AttnDev a{};
  a.q = p->q; a.k = p->k; a.v = p->v; a.o = p->o; a.lse = p->lse;
  a.q_bs = p->q_batch_stride; a.q_rs = p->q_row_stride; a.k_bs = p->k_batch_stride; a.k_rs = p->k_row_stride;
  a.v_bs = p->v_batch_stride; a.v_rs = p->v_row_stride; a.o_bs = p->o_batch_stride; a.o_rs = p->o_row_stride;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.dh = p->head_dim;
  a.scale = 1.0f / sqrtf((float)p->head_dim);
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
