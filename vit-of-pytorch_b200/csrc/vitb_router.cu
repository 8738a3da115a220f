// vitb_router.cu — Res-ViT routing kernels (HBM-bound, elementwise / small reductions).
//
//   vitb_router_decide_fwd/bwd : the decision tail of RouterModule.forward (res-vit/model.py:189-211):
//        2-way softmax, router entropy over the non-reserved tokens, hard keep/skip decision
//        (Gumbel-softmax straight-through in training, argmax in eval), reserved-token override and
//        the MSB-first bit-packing of _router2indices (:169-173) — one pass, one thread per token.
//   vitb_token_mean_fwd/bwd    : the global feature mean(x_embed[:, r0:], dim=1) (:180-184).
//   vitb_select_rows           : out[t,:] = member(index[t]) ? a[t,:] : b[t,:] with member() a 32-bit lookup
//        mask over the packed index — torch.isin (:469-472) + the train/eval blend (:487,:524) and, with
//        b = 0 and a one-bit mask, the row selection of BlockPathApproximators (:349-368) — no host sync.
#include "../../include/vitb200.h"
#include "vitb_common.cuh"

namespace {
using namespace vitb;

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
router_decide_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ noise, int T, int N, int bs,
                         int r0, int training, float tau, float* __restrict__ soft, float* __restrict__ hard,
                         float* __restrict__ ysoft, float* __restrict__ indices, float* __restrict__ entropy_sum) {
  __shared__ float s_part[kThreads / 32];
  float ent = 0.f;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int n = t % N;
    const bool reserved = n < r0;
    float packed = 0.f;
    for (int i = 0; i < bs; ++i) {
      const long long o = (static_cast<long long>(t) * bs + i) * 2;
      const float l0 = logits[o], l1 = logits[o + 1];
      const float m = fmaxf(l0, l1);
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      const float inv = 1.0f / (e0 + e1);
      const float p0 = e0 * inv, p1 = e1 * inv;
      soft[o] = p0;
      soft[o + 1] = p1;
      if (!reserved) ent -= p0 * logf(p0 + 1e-8f) + p1 * logf(p1 + 1e-8f);
      int keep;
      if (training) {
        const float a0 = (l0 + noise[o]) / tau, a1 = (l1 + noise[o + 1]) / tau;
        const float mm = fmaxf(a0, a1);
        const float f0 = expf(a0 - mm), f1 = expf(a1 - mm);
        const float iv = 1.0f / (f0 + f1);
        const float y0 = f0 * iv, y1 = f1 * iv;
        ysoft[o] = y0;
        ysoft[o + 1] = y1;
        keep = y1 > y0 ? 1 : 0;  // max() returns the first maximal index on ties
      } else {
        keep = p1 > p0 ? 1 : 0;  // argmax: first maximal index on ties
      }
      if (reserved) keep = 1;
      hard[o] = keep ? 0.f : 1.f;
      hard[o + 1] = keep ? 1.f : 0.f;
      packed += keep ? static_cast<float>(1 << (bs - 1 - i)) : 0.f;
    }
    indices[t] = packed;
  }
  ent = warp_sum(ent);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ent;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) tot += s_part[w];
    atomicAdd(entropy_sum, tot);
  }
}

// d_logits from: d_soft (gradient wrt softmax(logits)), d_entropy (scalar gradient wrt the normalised
// entropy), and in training d_hard flowing straight-through into softmax((logits+g)/tau).
__global__ void __launch_bounds__(kThreads)
router_decide_bwd_kernel(const float* __restrict__ soft, const float* __restrict__ ysoft,
                         const float* __restrict__ d_soft, const float* __restrict__ d_hard,
                         const float* __restrict__ d_entropy, float ent_scale, int T, int N, int bs, int r0,
                         int training, float tau, float* __restrict__ d_logits) {
  const float dent = d_entropy ? (*d_entropy) * ent_scale : 0.f;
  const long long total = static_cast<long long>(T) * bs;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(e / bs);
    const bool reserved = (t % N) < r0;
    const long long o = e * 2;
    const float p0 = soft[o], p1 = soft[o + 1];
    float g0 = d_soft ? d_soft[o] : 0.f, g1 = d_soft ? d_soft[o + 1] : 0.f;
    if (!reserved && dent != 0.f) {
      g0 -= dent * (logf(p0 + 1e-8f) + p0 / (p0 + 1e-8f));
      g1 -= dent * (logf(p1 + 1e-8f) + p1 / (p1 + 1e-8f));
    }
    const float dot = p0 * g0 + p1 * g1;
    float dl0 = p0 * (g0 - dot), dl1 = p1 * (g1 - dot);
    if (training && d_hard && !reserved) {
      const float y0 = ysoft[o], y1 = ysoft[o + 1];
      const float h0 = d_hard[o], h1 = d_hard[o + 1];
      const float dd = y0 * h0 + y1 * h1;
      dl0 += y0 * (h0 - dd) / tau;
      dl1 += y1 * (h1 - dd) / tau;
    }
    d_logits[o] = dl0;
    d_logits[o + 1] = dl1;
  }
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads)
token_mean_fwd_kernel(const void* __restrict__ x_, int Bsz, int N, int Cc, int r0, float* __restrict__ out) {
  const int total = Bsz * Cc;
  const float inv = 1.0f / static_cast<float>(N - r0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / Cc, c = i - b * Cc;
    float s = 0.f;
    for (int n = r0; n < N; ++n) {
      const long long o = (static_cast<long long>(b) * N + n) * Cc + c;
      s += BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x_)[o]) : reinterpret_cast<const float*>(x_)[o];
    }
    out[i] = s * inv;
  }
}

// dx[b, n, :] = dg[b, :] / (N - r0) for n >= r0, else 0.  A thread writes V consecutive columns of one token (16 bytes
// in bf16); Cc % V == 0 is required for the vector form (the scalar form takes the rest).  32-bit index arithmetic:
// the per-element 64-bit divisions of the first version made this broadcast cost 42 us for 13 M values.
template <bool BF16, int V>
__global__ void __launch_bounds__(kThreads)
token_mean_bwd_kernel(const float* __restrict__ dg, int Bsz, int N, int Cc, int r0, void* __restrict__ dx_) {
  const int cv = Cc / V;
  const long long total = static_cast<long long>(Bsz) * N * cv;
  const float inv = 1.0f / static_cast<float>(N - r0);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned bn = static_cast<unsigned>(i / cv);
    const int c = static_cast<int>(i - static_cast<long long>(bn) * cv) * V;
    const int b = static_cast<int>(bn / static_cast<unsigned>(N)), n = static_cast<int>(bn - static_cast<unsigned>(b) * N);
    float v[V];
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = n >= r0 ? __ldg(dg + b * Cc + c + j) * inv : 0.f;
    const long long o = static_cast<long long>(bn) * Cc + c;
    if constexpr (BF16) {
      if constexpr (V == 8) {
        uint4 w;
        w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]); w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dx_) + o) = w;
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) reinterpret_cast<__nv_bfloat16*>(dx_)[o + j] = __float2bfloat16(v[j]);
      }
    } else {
      if constexpr (V == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dx_) + o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) reinterpret_cast<float*>(dx_)[o + j] = v[j];
      }
    }
  }
}

// out[t, :] = ((mask >> int(index[t])) & 1) ? a[t, :] : b[t, :]   (null a / b read as zero)
template <bool BF16>
__global__ void __launch_bounds__(kThreads)
select_rows_kernel(const void* __restrict__ a_, const void* __restrict__ b_, const float* __restrict__ index,
                   unsigned mask, int rows, int cols, void* __restrict__ out_, int* __restrict__ any_member) {
  constexpr int V = BF16 ? 8 : 4;  // elements per 16-byte vector
  const int cv = cols / V;
  const long long total = static_cast<long long>(rows) * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / cv);
    const int idx = static_cast<int>(index[t]);
    const bool sel = (idx >= 0 && idx < 32) ? ((mask >> idx) & 1u) : false;
    if (any_member != nullptr && sel && i % cv == 0) *any_member = 1;   // benign race: every writer stores 1
    const void* src = sel ? a_ : b_;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (src) v = reinterpret_cast<const uint4*>(src)[i];
    reinterpret_cast<uint4*>(out_)[i] = v;
  }
}

inline int grid_for(long long items, int per_sm = 8) {
  long long blocks = (items + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(vitb_num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ---- Res-ViT scalar losses -----------------------------------------------------------------
// Reductions over 10^5 values that also write the gradient the backward needs, so the autograd node never launches a
// second pass.  They started as single-CTA kernels; at the benchmarked geometry that was 0.13 ms per DistillLoss call
// (ten per step) and 0.43 ms per ActiveLoss call — 4.7 % of the Res-ViT fine-tune step (profiles/launches_r02z_c5.csv) —
// latency-bound on one SM.  Now: a grid of blocks, one atomic per block.
constexpr int kLossThreads = 512;

__device__ __forceinline__ float block_sum(float v, float* s_part) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (kLossThreads >> 5) ? s_part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) s_part[0] = t;
  }
  __syncthreads();
  t = s_part[0];
  __syncthreads();
  return t;
}

// DistillLoss (res-vit/model.py:40-59): loss = mean((s - t)^2) over rows x cols values; ds = 2 (s - t) / (rows cols).
// s / t are [rows, cols] slices with row strides (the class-token rows student_out[:, 0, :] / teacher_out[:, 0, :]).
template <bool BF16>
__global__ void __launch_bounds__(kLossThreads)
distill_loss_kernel(const void* __restrict__ s_, long long s_stride, const void* __restrict__ t_, long long t_stride,
                    int rows, int cols, float* __restrict__ loss_acc, float* __restrict__ ds) {
  __shared__ float s_part[kLossThreads / 32];
  const long long n = static_cast<long long>(rows) * cols;
  const float inv = 1.0f / static_cast<float>(n);
  float acc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * kLossThreads;
  for (long long i = static_cast<long long>(blockIdx.x) * kLossThreads + threadIdx.x; i < n; i += stride) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i - static_cast<long long>(r) * cols);
    float a, b;
    if (BF16) {
      a = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s_)[r * s_stride + c]);
      b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(t_)[r * t_stride + c]);
    } else {
      a = reinterpret_cast<const float*>(s_)[r * s_stride + c];
      b = reinterpret_cast<const float*>(t_)[r * t_stride + c];
    }
    const float d = a - b;
    acc += d * d;
    if (ds) ds[i] = 2.f * d * inv;
  }
  const float tot = block_sum(acc, s_part);
  if (threadIdx.x == 0) atomicAdd(loss_acc, tot * inv);   // accumulates: d_loss sums over the dynamic layers (:648-650)
}

// ActiveLoss (res-vit/model.py:61-85): ratio = mean of p[b, n >= r0, j]; loss = (ratio + shift - target)^2;
// dp = 2 (ratio + shift - target) / count for n >= r0, else 0.  p is [B, N, L] fp32.  `shift` (device scalar, optional)
// carries global-batch mean minus this shard's mean under data parallelism (resvit.ActiveLoss.sync_group).
// Three launches: the sum over a grid of blocks into *sum (zeroed by a memset node), the gradient from it, the two scalars.
__global__ void __launch_bounds__(kLossThreads)
active_sum_kernel(const float* __restrict__ p, long long n, int N, int L, int r0, float* __restrict__ sum) {
  __shared__ float s_part[kLossThreads / 32];
  float acc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * kLossThreads;
  for (long long i = static_cast<long long>(blockIdx.x) * kLossThreads + threadIdx.x; i < n; i += stride) {
    const int tok = static_cast<int>((i / L) % N);
    if (tok >= r0) acc += p[i];
  }
  const float tot = block_sum(acc, s_part);
  if (threadIdx.x == 0) atomicAdd(sum, tot);
}

__global__ void __launch_bounds__(kLossThreads)
active_grad_kernel(const float* __restrict__ sum, long long n, int N, int L, int r0, float count, float target,
                   const float* __restrict__ shift, float* __restrict__ dp) {
  const float e = *sum / count + (shift ? *shift : 0.f) - target;
  const float g = 2.f * e / count;
  const long long stride = static_cast<long long>(gridDim.x) * kLossThreads;
  for (long long i = static_cast<long long>(blockIdx.x) * kLossThreads + threadIdx.x; i < n; i += stride) {
    const int tok = static_cast<int>((i / L) % N);
    dp[i] = tok >= r0 ? g : 0.f;
  }
}

__global__ void active_final_kernel(float* __restrict__ sum_then_ratio, float count, float target, const float* __restrict__ shift,
                                    float* __restrict__ loss) {
  const float ratio = *sum_then_ratio / count;
  const float e = ratio + (shift ? *shift : 0.f) - target;
  *sum_then_ratio = ratio;
  if (loss) *loss = e * e;
}

}  // namespace

extern "C" {

int vitb_router_decide_fwd(const float* logits, const float* noise, int T, int N, int block_size,
                           int reserve_initials, int training, float tau, float* soft, float* hard,
                           float* ysoft, float* indices, float* entropy_sum, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (T == 0) return VITB_OK;
  VITB_REQUIRE(logits && soft && hard && indices && entropy_sum && T > 0 && N > 0, VITB_ERR_BAD_ARG,
               "router_decide_fwd: bad args");
  VITB_REQUIRE(block_size >= 1 && block_size <= 5, VITB_ERR_UNSUPPORTED_SHAPE, "router_decide_fwd: block_size %d", block_size);
  VITB_REQUIRE(!training || (noise && ysoft), VITB_ERR_BAD_ARG, "router_decide_fwd: training needs noise and ysoft");
  router_decide_fwd_kernel<<<grid_for(T, 2), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      logits, noise, T, N, block_size, reserve_initials, training, tau, soft, hard, ysoft, indices, entropy_sum);
  VITB_LAUNCH_CHECK("router_decide_fwd_kernel");
  return VITB_OK;
}

int vitb_router_decide_bwd(const float* soft, const float* ysoft, const float* d_soft, const float* d_hard,
                           const float* d_entropy, float entropy_scale, int T, int N, int block_size,
                           int reserve_initials, int training, float tau, float* d_logits, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (T == 0) return VITB_OK;
  VITB_REQUIRE(soft && d_logits && T > 0 && N > 0, VITB_ERR_BAD_ARG, "router_decide_bwd: bad args");
  router_decide_bwd_kernel<<<grid_for(static_cast<long long>(T) * block_size, 2), kThreads, 0,
                             reinterpret_cast<cudaStream_t>(stream_)>>>(
      soft, ysoft, d_soft, d_hard, d_entropy, entropy_scale, T, N, block_size, reserve_initials, training, tau, d_logits);
  VITB_LAUNCH_CHECK("router_decide_bwd_kernel");
  return VITB_OK;
}

int vitb_token_mean_fwd(const void* x, int dtype, int B, int N, int C, int reserve_initials, float* out, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (B == 0) return VITB_OK;
  VITB_REQUIRE(x && out && N > reserve_initials && C > 0, VITB_ERR_BAD_ARG, "token_mean_fwd: bad args");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  if (dtype == VITB_BF16) token_mean_fwd_kernel<true><<<grid_for(static_cast<long long>(B) * C), kThreads, 0, s>>>(x, B, N, C, reserve_initials, out);
  else token_mean_fwd_kernel<false><<<grid_for(static_cast<long long>(B) * C), kThreads, 0, s>>>(x, B, N, C, reserve_initials, out);
  VITB_LAUNCH_CHECK("token_mean_fwd_kernel");
  return VITB_OK;
}

int vitb_token_mean_bwd(const float* dg, int dtype, int B, int N, int C, int reserve_initials, void* dx, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (B == 0) return VITB_OK;
  VITB_REQUIRE(dg && dx && N > reserve_initials && C > 0, VITB_ERR_BAD_ARG, "token_mean_bwd: bad args");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = static_cast<long long>(B) * N * C;
  const bool al16 = (reinterpret_cast<uintptr_t>(dx) & 15u) == 0;
  if (dtype == VITB_BF16) {
    if (C % 8 == 0 && al16) token_mean_bwd_kernel<true, 8><<<grid_for(total / 8), kThreads, 0, s>>>(dg, B, N, C, reserve_initials, dx);
    else token_mean_bwd_kernel<true, 1><<<grid_for(total), kThreads, 0, s>>>(dg, B, N, C, reserve_initials, dx);
  } else {
    if (C % 4 == 0 && al16) token_mean_bwd_kernel<false, 4><<<grid_for(total / 4), kThreads, 0, s>>>(dg, B, N, C, reserve_initials, dx);
    else token_mean_bwd_kernel<false, 1><<<grid_for(total), kThreads, 0, s>>>(dg, B, N, C, reserve_initials, dx);
  }
  VITB_LAUNCH_CHECK("token_mean_bwd_kernel");
  return VITB_OK;
}

int vitb_select_rows_flag(const void* a, const void* b, const float* index, uint32_t member_mask, int rows, int cols,
                          int dtype, void* out, int32_t* any_member, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (rows == 0 || cols == 0) return VITB_OK;
  VITB_REQUIRE(index && out && rows > 0, VITB_ERR_BAD_ARG, "select_rows: bad args");
  const int V = dtype == VITB_BF16 ? 8 : 4;
  VITB_REQUIRE(cols % V == 0, VITB_ERR_UNSUPPORTED_SHAPE, "select_rows: cols %d must be a multiple of %d", cols, V);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = static_cast<long long>(rows) * (cols / V);
  if (dtype == VITB_BF16) select_rows_kernel<true><<<grid_for(total), kThreads, 0, s>>>(a, b, index, member_mask, rows, cols, out, any_member);
  else select_rows_kernel<false><<<grid_for(total), kThreads, 0, s>>>(a, b, index, member_mask, rows, cols, out, any_member);
  VITB_LAUNCH_CHECK("select_rows_kernel");
  return VITB_OK;
}

int vitb_select_rows(const void* a, const void* b, const float* index, uint32_t member_mask, int rows, int cols,
                     int dtype, void* out, void* stream_) {
  return vitb_select_rows_flag(a, b, index, member_mask, rows, cols, dtype, out, nullptr, stream_);
}

int vitb_distill_loss(const void* student, int64_t s_row_stride, const void* teacher, int64_t t_row_stride, int dtype, int rows,
                      int cols, float* loss_acc, float* d_student, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(student && teacher && loss_acc && rows > 0 && cols > 0, VITB_ERR_BAD_ARG, "distill_loss: bad args");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = grid_for((static_cast<long long>(rows) * cols + 3) / 4, 2);   // about four values per thread
  if (dtype == VITB_BF16) distill_loss_kernel<true><<<grid, kLossThreads, 0, s>>>(student, s_row_stride, teacher, t_row_stride, rows, cols, loss_acc, d_student);
  else distill_loss_kernel<false><<<grid, kLossThreads, 0, s>>>(student, s_row_stride, teacher, t_row_stride, rows, cols, loss_acc, d_student);
  VITB_LAUNCH_CHECK("distill_loss_kernel");
  return VITB_OK;
}

int vitb_active_loss(const float* probs, int B, int N, int L, int reserve_initials, float target, const float* shift_dev,
                     float* ratio_out, float* loss, float* d_probs, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(probs && ratio_out && B > 0 && N > reserve_initials && L > 0, VITB_ERR_BAD_ARG,
               "active_loss: bad args (ratio_out is required: it carries the sum between the launches)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  const long long n = static_cast<long long>(B) * N * L;
  const float count = static_cast<float>(static_cast<long long>(B) * (N - reserve_initials) * L);
  const int grid = grid_for((n + 3) / 4, 2);
  VITB_CUDA_CHECK(cudaMemsetAsync(ratio_out, 0, sizeof(float), s));
  active_sum_kernel<<<grid, kLossThreads, 0, s>>>(probs, n, N, L, reserve_initials, ratio_out);
  if (d_probs) active_grad_kernel<<<grid, kLossThreads, 0, s>>>(ratio_out, n, N, L, reserve_initials, count, target, shift_dev, d_probs);
  active_final_kernel<<<1, 1, 0, s>>>(ratio_out, count, target, shift_dev, loss);
  VITB_LAUNCH_CHECK("active_loss_kernel");
  return VITB_OK;
}

}  // extern "C"
