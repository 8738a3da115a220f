// vitb_elementwise.cu — the HBM-bound, non-contraction kernels of the ViT encoder path:
// operand casts / bf16 hi-lo split, patch extraction, class-token + position rows, the embedding
// backward reduction, column sums (bias gradients), cross-entropy, fused SGD / AdamW.
// All use 128-bit coalesced access and grid sizes in multiples of the SM count.
#include <stdlib.h>

#include "../../include/vitb200.h"
#include "vitb_common.cuh"

namespace {
using namespace vitb;

constexpr int kThreads = 256;

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

inline int grid_for(long long work_items, int per_sm = 8) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  long long cap = static_cast<long long>(vitb_num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 cast, optionally with the residual "lo" half (x - bf16(x)) for the bf16x3 GEMMs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
cast_split_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo) {
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 h;
    h.x = pack_bf16x2(v.x, v.y);
    h.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(hi)[i] = h;
    if (lo) {
      uint2 l;
      l.x = pack_bf16x2(v.x - bf16_round(v.x), v.y - bf16_round(v.y));
      l.y = pack_bf16x2(v.z - bf16_round(v.z), v.w - bf16_round(v.w));
      reinterpret_cast<uint2*>(lo)[i] = l;
    }
  }
  // tail
  const long long t0 = n4 << 2;
  for (long long i = t0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    hi[i] = __float2bfloat16(v);
    if (lo) lo[i] = __float2bfloat16(v - bf16_round(v));
  }
}

// ---------------------------------------------------------------------------------------------
// Patch extraction (the Conv2d(3,D,P,P) of src/model.py:179,197 as a GEMM operand):
// rows = (b, py, px), k = (c, ph, pw) — the flattening of conv weight [D,3,P,P]; K padded to ldk.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
im2col_kernel(const float* __restrict__ img, int Bsz, int Cin, int H, int W, int P, int gh, int gw,
              int ldk, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int kgroups = ldk >> 3;
  const long long total = static_cast<long long>(Bsz) * gh * gw * kgroups;
  const int Kreal = Cin * P * P;
  // 16-byte loads need every group start 16-byte aligned: patch width a multiple of 8 pixels, row pitch a multiple of
  // 4 floats, aligned base
  const bool vec8 = (P % 8 == 0) && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15u) == 0);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int kg = static_cast<int>(i % kgroups);
    const long long r = i / kgroups;
    const int px = static_cast<int>(r % gw);
    const int py = static_cast<int>((r / gw) % gh);
    const int b = static_cast<int>(r / (static_cast<long long>(gw) * gh));
    float v[8];
    if (vec8 && kg * 8 + 8 <= Kreal) {
      // P % 8 == 0: the eight k of this group are eight consecutive pixels of one patch row -> two 16-byte loads
      const int k = kg * 8;
      const int c = k / (P * P);
      const int rem = k - c * P * P;
      const int ph = rem / P, pw = rem - ph * P;
      const float4* src = reinterpret_cast<const float4*>(
          img + ((static_cast<long long>(b) * Cin + c) * H + (py * P + ph)) * W + (px * P + pw));
      const float4 a0 = __ldg(src), a1 = __ldg(src + 1);
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kg * 8 + j;
        float val = 0.f;
        if (k < Kreal) {
          const int c = k / (P * P);
          const int rem = k - c * P * P;
          const int ph = rem / P, pw = rem - ph * P;
          val = img[((static_cast<long long>(b) * Cin + c) * H + (py * P + ph)) * W + (px * P + pw)];
        }
        v[j] = val;
      }
    }
    uint4 h;
    h.x = pack_bf16x2(v[0], v[1]); h.y = pack_bf16x2(v[2], v[3]);
    h.z = pack_bf16x2(v[4], v[5]); h.w = pack_bf16x2(v[6], v[7]);
    reinterpret_cast<uint4*>(hi + r * ldk)[kg] = h;
    if (lo) {
      uint4 l;
      l.x = pack_bf16x2(v[0] - bf16_round(v[0]), v[1] - bf16_round(v[1]));
      l.y = pack_bf16x2(v[2] - bf16_round(v[2]), v[3] - bf16_round(v[3]));
      l.z = pack_bf16x2(v[4] - bf16_round(v[4]), v[5] - bf16_round(v[5]));
      l.w = pack_bf16x2(v[6] - bf16_round(v[6]), v[7] - bf16_round(v[7]));
      reinterpret_cast<uint4*>(lo + r * ldk)[kg] = l;
    }
  }
}

// x[b, 0, :] = cls + pos[0]   (torch.cat([cls_token.repeat], ...) + pos_embedding, src/model.py:203-204,17)
__global__ void __launch_bounds__(kThreads)
cls_rows_kernel(float* __restrict__ x, int Bsz, int N, int D, const float* __restrict__ cls,
                const float* __restrict__ pos) {
  const int total = Bsz * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / D, d = i - b * D;
    x[static_cast<long long>(b) * N * D + d] = cls[d] + (pos ? pos[d] : 0.f);
  }
}

// Backward of the embedding stage: dx is the gradient of (cat(cls, patches) + pos) [B,N,D].
//   dpos[n,d]  += sum_b dx[b,n,d]                   (pos_embedding grad)
//   dcls[d]    += sum_b dx[b,0,d]                   (cls_token grad)
//   dbias[d]   += sum_{b,n>=1} dx[b,n,d]            (conv bias grad)
//   dpatch[b*np + n-1, d] = bf16(dx[b,n,d]) (+ lo)  (A^T operand of the conv-weight wgrad GEMM)
__global__ void __launch_bounds__(kThreads)
embed_bwd_kernel(const float* __restrict__ dx, int Bsz, int N, int D, float* __restrict__ dpos,
                 float* __restrict__ dcls, float* __restrict__ dbias, __nv_bfloat16* __restrict__ dp_hi,
                 __nv_bfloat16* __restrict__ dp_lo) {
  const int d4 = D >> 2;
  const int total = N * d4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / d4, c = (i - n * d4) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < Bsz; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(dx + (static_cast<long long>(b) * N + n) * D + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      if (n >= 1 && dp_hi) {
        const long long off = (static_cast<long long>(b) * (N - 1) + (n - 1)) * D + c;
        uint2 h;
        h.x = pack_bf16x2(v.x, v.y); h.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dp_hi + off) = h;
        if (dp_lo) {
          uint2 l;
          l.x = pack_bf16x2(v.x - bf16_round(v.x), v.y - bf16_round(v.y));
          l.y = pack_bf16x2(v.z - bf16_round(v.z), v.w - bf16_round(v.w));
          *reinterpret_cast<uint2*>(dp_lo + off) = l;
        }
      }
    }
    if (dpos) {
      float* p = dpos + static_cast<long long>(n) * D + c;
      atomicAdd(p + 0, s.x); atomicAdd(p + 1, s.y); atomicAdd(p + 2, s.z); atomicAdd(p + 3, s.w);
    }
    if (n == 0 && dcls) {
      atomicAdd(dcls + c + 0, s.x); atomicAdd(dcls + c + 1, s.y);
      atomicAdd(dcls + c + 2, s.z); atomicAdd(dcls + c + 3, s.w);
    }
    if (n >= 1 && dbias) {
      atomicAdd(dbias + c + 0, s.x); atomicAdd(dbias + c + 1, s.y);
      atomicAdd(dbias + c + 2, s.z); atomicAdd(dbias + c + 3, s.w);
    }
  }
}

// out[c] += sum_r x[r, c].  A thread owns 16 bytes of a row (8 bf16 / 4 fp32 columns; 4 bf16 columns when the width
// is not a multiple of 8).  Block = (column groups) x (NY row lanes); the block's row slice is walked NY rows at a
// time with 8 independent loads per thread, the NY partial sums meet in shared memory and one 16-byte vector reduction
// per four columns and block goes out.  HBM-bound: rows * cols * esize bytes read once; ptxas keeps the kernel at 32
// registers, so the loads in flight come from occupancy; the grid is sized for 32 warps per SM (measured best).
constexpr int kColsumThreads = 512;
template <bool BF16, int VEC>   // VEC = columns per thread: 8 or 4 (bf16), 4 (fp32)
__global__ void __launch_bounds__(kColsumThreads)
colsum_kernel(const void* __restrict__ x_, int rows, int cols, long long ld_, float* __restrict__ out0,
              float* __restrict__ out1, float* __restrict__ out2, int seg_cols, bool v4) {
  extern __shared__ __align__(16) float colsum_part[];   // [blockDim.y][blockDim.x * VEC]
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const bool live = c < cols;
  const int rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per;
  const int r1 = min(rows, r0 + rows_per);
  const int ny = blockDim.y;
  float s[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s[j] = 0.f;
  constexpr int ESIZE = BF16 ? 2 : 4;
  const char* base = reinterpret_cast<const char*>(x_) + static_cast<long long>(c) * ESIZE;
  const long long pitch = ld_ * ESIZE;
  auto acc = [&](const uint4& u) {
    if constexpr (!BF16) {
      s[0] += __uint_as_float(u.x); s[1] += __uint_as_float(u.y); s[2] += __uint_as_float(u.z); s[3] += __uint_as_float(u.w);
    } else {
      s[0] += bf16_lo(u.x); s[1] += bf16_hi(u.x); s[2] += bf16_lo(u.y); s[3] += bf16_hi(u.y);
      if constexpr (VEC == 8) { s[4] += bf16_lo(u.z); s[5] += bf16_hi(u.z); s[6] += bf16_lo(u.w); s[7] += bf16_hi(u.w); }
    }
  };
  auto ld = [&](int r) {
    const char* p = base + static_cast<long long>(r) * pitch;
    if constexpr (BF16 && VEC == 4) {
      const uint2 u = *reinterpret_cast<const uint2*>(p);
      return make_uint4(u.x, u.y, 0u, 0u);
    } else {
      return *reinterpret_cast<const uint4*>(p);
    }
  };
  if (live) {
    int r = r0 + threadIdx.y;
    for (; r + 7 * ny < r1; r += 8 * ny) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ld(r + j * ny);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc(v[j]);
    }
    for (; r < r1; r += ny) acc(ld(r));
  }
  const int width = blockDim.x * VEC;
#pragma unroll
  for (int j = 0; j < VEC; ++j) colsum_part[threadIdx.y * width + threadIdx.x * VEC + j] = s[j];
  __syncthreads();
  // fold the row lanes; consecutive threads own consecutive columns, so a warp's 32 scalar reductions are one 128-byte
  // transaction at L2.  (The first version let the lanes that loaded 8 columns each reduce them: 32 bytes between lanes,
  // eight transactions per warp instruction, issued by a fifth of the block — the tail was the longest part of the
  // kernel, profiles/colsum_r02.txt.  VITB_COLSUM_V4=1 sends red.global.add.v4.f32 per four columns instead, for A/B runs.)
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthr = blockDim.x * blockDim.y;
  const int cbase = blockIdx.x * width;
  if (!v4) {
    for (int cl = tid; cl < width; cl += nthr) {
      const int cc = cbase + cl;
      if (cc >= cols) break;
      float t = colsum_part[cl];
      for (int y = 1; y < ny; ++y) t += colsum_part[y * width + cl];
      const int seg = cc / seg_cols;
      float* dst = (seg == 0 ? out0 : (seg == 1 ? out1 : out2)) - static_cast<long long>(seg) * seg_cols + cc;
      atomicAdd(dst, t);
    }
    return;
  }
  for (int q4 = tid; q4 * 4 < width; q4 += nthr) {
    const int cc = cbase + q4 * 4;
    if (cc >= cols) break;
    float4 t = *reinterpret_cast<const float4*>(colsum_part + q4 * 4);
    for (int y = 1; y < ny; ++y) {
      const float4 u = *reinterpret_cast<const float4*>(colsum_part + y * width + q4 * 4);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const int seg = cc / seg_cols;    // seg_cols is a multiple of 4: a quad never straddles two segments
    float* dst = (seg == 0 ? out0 : (seg == 1 ? out1 : out2)) - static_cast<long long>(seg) * seg_cols + cc;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
    } else {
      atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
    }
  }
}

// grid for the above: x over column groups with the block width that wastes the fewest lanes (2304 bf16 columns are
// 288 groups = 3 blocks of 96 x 5), y over row slices of >= 8 rows per row lane, 32 warps per SM in total
template <bool BF16, int VEC>
static void colsum_launch(const void* x, int rows, int cols, long long ld, float* o0, float* o1, float* o2, int seg_cols,
                          cudaStream_t stream) {
  const int groups = cols / VEC;
  int tb = 32, waste = 1 << 30;
  for (int cand = 256; cand >= 32; cand -= 32) {
    const int w = (groups + cand - 1) / cand * cand - groups;
    if (w < waste) { waste = w; tb = cand; }
  }
  const int ny = kColsumThreads / tb;                 // >= 2 row lanes share a block (and one atomic per column)
  const int bx = (groups + tb - 1) / tb;
  const int warps_per_block = tb * ny / 32;
  int warps_per_sm = 32;     // measured best of 4..64 on B200 (profiles/colsum_r02.txt)
  if (const char* e = getenv("VITB_COLSUM_WARPS")) { const int v = atoi(e); if (v >= 1 && v <= 64) warps_per_sm = v; }   // tuning runs
  int by = (vitb_num_sms() * warps_per_sm + bx * warps_per_block - 1) / (bx * warps_per_block);
  const int max_by = (rows + 8 * ny - 1) / (8 * ny);     // at least one round of 8 loads per thread
  if (by > max_by) by = max_by;
  if (by < 1) by = 1;
  const size_t smem = static_cast<size_t>(ny) * tb * VEC * sizeof(float);
  bool v4 = false;
  if (const char* e = getenv("VITB_COLSUM_V4")) v4 = atoi(e) != 0;     // tuning runs
  colsum_kernel<BF16, VEC><<<dim3(bx, by), dim3(tb, ny), smem, stream>>>(x, rows, cols, ld, o0, o1, o2, seg_cols, v4);
}
static void colsum_dispatch(const void* x, int x_dtype, int rows, int cols, long long ld, float* o0, float* o1, float* o2,
                            int seg_cols, cudaStream_t stream) {
  if (x_dtype == VITB_BF16) {
    const bool v8 = seg_cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
    if (v8) colsum_launch<true, 8>(x, rows, cols, ld, o0, o1, o2, seg_cols, stream);
    else colsum_launch<true, 4>(x, rows, cols, ld, o0, o1, o2, seg_cols, stream);
  } else {
    colsum_launch<false, 4>(x, rows, cols, ld, o0, o1, o2, seg_cols, stream);
  }
}

// Cross entropy (mean) forward + gradient in one block: nn.CrossEntropyLoss, src/train.py:151,22;
// res-vit/model.py:550,681.  loss = mean_b(lse_b - logit[b,label_b]); dlogits = (softmax - onehot)/B.
__global__ void __launch_bounds__(1024)
ce_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int Bsz, int Ccls,
          float* __restrict__ loss, float* __restrict__ dlogits) {
  __shared__ float s_part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int b = warp; b < Bsz; b += 32) {
    const float* row = logits + static_cast<long long>(b) * Ccls;
    float mx = -INFINITY;
    for (int c = lane; c < Ccls; c += 32) mx = fmaxf(mx, row[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < Ccls; c += 32) se += expf(row[c] - mx);
    se = warp_sum(se);
    const int lab = static_cast<int>(labels[b]);
    const float lse = mx + logf(se);
    if (lane == 0) acc += lse - row[lab];
    if (dlogits) {
      const float inv = 1.0f / se, sc = 1.0f / Bsz;
      for (int c = lane; c < Ccls; c += 32) {
        const float pr = expf(row[c] - mx) * inv;
        dlogits[static_cast<long long>(b) * Ccls + c] = (pr - (c == lab ? 1.f : 0.f)) * sc;
      }
    }
  }
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += s_part[w];
    *loss = t / Bsz;
  }
}

// SGD with momentum (torch.optim.SGD semantics, src/train.py:154-158) over a flat parameter buffer,
// refreshing the bf16 GEMM shadow (and its lo half in fp32 parity mode) in the same pass.
__global__ void __launch_bounds__(kThreads)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, long long n,
           float lr_host, const float* __restrict__ hyper_dev, float momentum, float dampening, float wd, int nesterov,
           int first_step, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  // device-resident hyper-parameters {lr, momentum, dampening, weight_decay} keep a captured CUDA graph following the
  // LR scheduler (OneCycleLR cycles the momentum as well as the learning rate)
  const float lr = hyper_dev ? hyper_dev[0] : lr_host;
  if (hyper_dev) { momentum = hyper_dev[1]; dampening = hyper_dev[2]; wd = hyper_dev[3]; }
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv0 = reinterpret_cast<const float4*>(g)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w};
    float ga[4] = {gv0.x, gv0.y, gv0.z, gv0.w};
    float ma[4] = {0.f, 0.f, 0.f, 0.f};
    if (m && !first_step) {
      const float4 mv = reinterpret_cast<float4*>(m)[i];
      ma[0] = mv.x; ma[1] = mv.y; ma[2] = mv.z; ma[3] = mv.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = ga[j] + wd * pa[j];
      if (m) {
        ma[j] = first_step ? gg : momentum * ma[j] + (1.f - dampening) * gg;
        gg = nesterov ? gg + momentum * ma[j] : ma[j];
      }
      pa[j] -= lr * gg;
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    if (m) reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    if (hi) {
      uint2 h;
      h.x = pack_bf16x2(pa[0], pa[1]); h.y = pack_bf16x2(pa[2], pa[3]);
      reinterpret_cast<uint2*>(hi)[i] = h;
    }
    if (lo) {
      uint2 l;
      l.x = pack_bf16x2(pa[0] - bf16_round(pa[0]), pa[1] - bf16_round(pa[1]));
      l.y = pack_bf16x2(pa[2] - bf16_round(pa[2]), pa[3] - bf16_round(pa[3]));
      reinterpret_cast<uint2*>(lo)[i] = l;
    }
  }
}

// AdamW (torch.optim.AdamW semantics, res-vit/train.py:272-277); grad_scale_dev (optional, device)
// carries the clip_grad_norm_ coefficient (res-vit/train.py:65) so no host sync is needed.
__global__ void __launch_bounds__(kThreads)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, long long n, float lr_host, const float* __restrict__ hyper_dev, float b1, float b2,
             float eps, float wd, int step_host, const int* __restrict__ step_dev,
             const float* __restrict__ grad_scale_dev, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  // {lr, beta1, beta2, weight_decay} and the step counter may live on the device so that a captured CUDA graph keeps
  // following the scheduler (OneCycleLR cycles beta1)
  const float lr = hyper_dev ? hyper_dev[0] : lr_host;
  if (hyper_dev) { b1 = hyper_dev[1]; b2 = hyper_dev[2]; wd = hyper_dev[3]; }
  const float stepf = static_cast<float>(step_dev ? *step_dev : step_host);
  const float bc1 = 1.f - powf(b1, stepf), bc2 = 1.f - powf(b2, stepf);
  const float gs = grad_scale_dev ? *grad_scale_dev : 1.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pv = p[i];
    const float gv = g[i] * gs;
    pv *= (1.f - lr * wd);
    const float mv = b1 * m[i] + (1.f - b1) * gv;
    const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / sqrtf(bc2) + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv;
    if (hi) hi[i] = __float2bfloat16(pv);
    if (lo) lo[i] = __float2bfloat16(pv - bf16_round(pv));
  }
}

// The same update over a flat buffer whose parameters (segments) can be SKIPPED individually: torch.optim.AdamW leaves a
// parameter whose .grad is None untouched — no weight decay, no moment decay, its own step count does not advance
// (res-vit/train.py:272-277 with BlockPathApproximators, res-vit/model.py:349-368: an approximator whose key did not occur
// in the batch never ran).  seg_end[s] = end offset of segment s (ascending); seg_flag[s] = index into flags[] (int32,
// non-zero = the parameter received a gradient this step) or -1 = always live; seg_step[s] = that parameter's own step
// count (already advanced for this step by adamw_advance_kernel).  A block walks a contiguous chunk, so a thread finds its
// segment once (binary search) and then only moves a cursor.
__global__ void __launch_bounds__(kThreads)
adamw_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, const float* __restrict__ hyper_dev, float eps, const float* __restrict__ grad_scale_dev,
                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, const long long* __restrict__ seg_end,
                 const int* __restrict__ seg_flag, const int* __restrict__ flags, const float* __restrict__ seg_step,
                 int nseg) {
  const float lr = hyper_dev[0], b1 = hyper_dev[1], b2 = hyper_dev[2], wd = hyper_dev[3];
  const float gs = grad_scale_dev ? *grad_scale_dev : 1.f;
  const long long per_block = ((n + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
  const long long c0 = static_cast<long long>(blockIdx.x) * per_block;
  const long long c1 = c0 + per_block < n ? c0 + per_block : n;
  long long i = c0 + threadIdx.x;
  if (i >= c1) return;
  int lo_s = 0, hi_s = nseg - 1;
  while (lo_s < hi_s) {                       // first segment that ends behind i
    const int mid = (lo_s + hi_s) >> 1;
    if (seg_end[mid] > i) hi_s = mid; else lo_s = mid + 1;
  }
  int seg = lo_s;
  long long end = seg_end[seg];
  bool live = false;
  float bc1 = 1.f, bc2s = 1.f;
  auto load_seg = [&]() {
    const int f = seg_flag[seg];
    live = f < 0 || flags[f] != 0;
    const float stepf = seg_step[seg];
    bc1 = 1.f - powf(b1, stepf);
    bc2s = sqrtf(1.f - powf(b2, stepf));
  };
  load_seg();
  for (; i < c1; i += blockDim.x) {
    while (i >= end && seg < nseg - 1) { ++seg; end = seg_end[seg]; load_seg(); }
    if (!live) continue;
    float pv = p[i];
    const float gv = g[i] * gs;
    pv *= (1.f - lr * wd);
    const float mv = b1 * m[i] + (1.f - b1) * gv;
    const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2s + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv;
    if (hi) hi[i] = __float2bfloat16(pv);
    if (lo) lo[i] = __float2bfloat16(pv - bf16_round(pv));
  }
}

// seg_step[s] += 1 for every segment that is live this step
__global__ void adamw_advance_kernel(const int* __restrict__ seg_flag, const int* __restrict__ flags,
                                     float* __restrict__ seg_step, int nseg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) {
    const int f = seg_flag[s];
    if (f < 0 || flags[f] != 0) seg_step[s] += 1.f;
  }
}

// out[0] += sum x^2   (for clip_grad_norm_)
__global__ void __launch_bounds__(kThreads)
sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float s_part[kThreads / 32];
  float acc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += s_part[w];
    atomicAdd(out, t);
  }
}

// coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))     (torch clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef,
                                 float* __restrict__ norm_out) {
  const float nrm = sqrtf(*sumsq);
  if (norm_out) *norm_out = nrm;
  const float c = max_norm / (nrm + 1e-6f);
  *coef = c < 1.f ? c : 1.f;
}

// dz = dy * gelu'(z)  (exact-erf GELU, nn.GELU() of src/model.py:33 / res-vit/model.py:154,158,160,312)
template <bool BF16>
__global__ void __launch_bounds__(kThreads)
gelu_bwd_kernel(const void* __restrict__ dy_, const void* __restrict__ z_, void* __restrict__ out_, long long n) {
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 d, z;
    if constexpr (BF16) {
      const uint2 a = reinterpret_cast<const uint2*>(dy_)[i], b = reinterpret_cast<const uint2*>(z_)[i];
      d = make_float4(bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y));
      z = make_float4(bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y));
    } else {
      d = reinterpret_cast<const float4*>(dy_)[i];
      z = reinterpret_cast<const float4*>(z_)[i];
    }
    d.x *= gelu_erf_grad(z.x); d.y *= gelu_erf_grad(z.y); d.z *= gelu_erf_grad(z.z); d.w *= gelu_erf_grad(z.w);
    if constexpr (BF16) {
      uint2 o;
      o.x = pack_bf16x2(d.x, d.y); o.y = pack_bf16x2(d.z, d.w);
      reinterpret_cast<uint2*>(out_)[i] = o;
    } else {
      reinterpret_cast<float4*>(out_)[i] = d;
    }
  }
  for (long long i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    if constexpr (BF16) {
      const float d = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dy_)[i]);
      const float z = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(z_)[i]);
      reinterpret_cast<__nv_bfloat16*>(out_)[i] = __float2bfloat16(d * gelu_erf_grad(z));
    } else {
      reinterpret_cast<float*>(out_)[i] =
          reinterpret_cast<const float*>(dy_)[i] * gelu_erf_grad(reinterpret_cast<const float*>(z_)[i]);
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Dropout (nn.Dropout of PositionEmbs / MlpBlock / EncoderBlock, src/model.py:11-20,34-50,110-123).
// y = [residual +] keep * x / (1 - p), keep ~ Bernoulli(1 - p) from Philox4x32-10 keyed by (seed, draw counter), one
// counter-mode call per four consecutive elements.  The draw counter lives in DEVICE memory (state[0]) and is advanced
// by the kernel itself, so a step replayed from a CUDA graph draws a fresh mask every replay: every block reads the
// counter when it starts; the block that takes the last ticket (state[1]) — by then every block has started — stores
// counter + 1 and clears the tickets.  The mask is kept (one byte per element) for the backward.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint4 c) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

template <bool BF16, bool RES>
__global__ void __launch_bounds__(kThreads)
dropout_fwd_kernel(const void* __restrict__ x_, const float* __restrict__ res, void* __restrict__ y_,
                   uint8_t* __restrict__ mask, long long n, uint32_t keep_below, float scale,
                   unsigned long long seed, unsigned long long* __restrict__ state) {
  const unsigned long long draw = *reinterpret_cast<volatile unsigned long long*>(state);
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const long long n4 = (n + 3) >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 r = philox4x32_10(k0, k1, make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32),
                                                     static_cast<uint32_t>(draw), static_cast<uint32_t>(draw >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    const long long e0 = i << 2;
    if (e0 + 3 < n) {
      float4 v;
      if constexpr (BF16) {
        const uint2 a = reinterpret_cast<const uint2*>(x_)[i];
        v = make_float4(bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y));
      } else {
        v = reinterpret_cast<const float4*>(x_)[i];
      }
      uchar4 m;
      m.x = rr[0] < keep_below; m.y = rr[1] < keep_below; m.z = rr[2] < keep_below; m.w = rr[3] < keep_below;
      v.x = m.x ? v.x * scale : 0.f; v.y = m.y ? v.y * scale : 0.f; v.z = m.z ? v.z * scale : 0.f; v.w = m.w ? v.w * scale : 0.f;
      reinterpret_cast<uchar4*>(mask)[i] = m;
      if constexpr (RES) {
        const float4 q = reinterpret_cast<const float4*>(res)[i];
        reinterpret_cast<float4*>(y_)[i] = make_float4(q.x + v.x, q.y + v.y, q.z + v.z, q.w + v.w);
      } else if constexpr (BF16) {
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
        reinterpret_cast<uint2*>(y_)[i] = o;
      } else {
        reinterpret_cast<float4*>(y_)[i] = v;
      }
    } else {
      for (int j = 0; j < 4 && e0 + j < n; ++j) {
        const long long e = e0 + j;
        float v = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x_)[e]) : reinterpret_cast<const float*>(x_)[e];
        const uint8_t m = rr[j] < keep_below;
        v = m ? v * scale : 0.f;
        mask[e] = m;
        if constexpr (RES) reinterpret_cast<float*>(y_)[e] = res[e] + v;
        else if constexpr (BF16) reinterpret_cast<__nv_bfloat16*>(y_)[e] = __float2bfloat16(v);
        else reinterpret_cast<float*>(y_)[e] = v;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(state + 1, 1ull);
    if (t == gridDim.x - 1) {       // every block has read `draw` before it took its ticket
      state[1] = 0ull;
      state[0] = draw + 1ull;
      __threadfence();
    }
  }
}

// dx = dy * mask / (1 - p)
template <bool IN_BF16, bool OUT_BF16>
__global__ void __launch_bounds__(kThreads)
dropout_bwd_kernel(const void* __restrict__ dy_, const uint8_t* __restrict__ mask, void* __restrict__ dx_, long long n,
                   float scale) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride) {
    const float d = IN_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dy_)[e]) : reinterpret_cast<const float*>(dy_)[e];
    const float v = mask[e] ? d * scale : 0.f;
    if constexpr (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(dx_)[e] = __float2bfloat16(v);
    else reinterpret_cast<float*>(dx_)[e] = v;
  }
}

extern "C" {

int vitb_gelu_bwd(const void* dy, const void* z, void* out, int64_t n, int dtype, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(dy && z && out && n > 0, VITB_ERR_BAD_ARG, "gelu_bwd: bad args");
  if (dtype == VITB_BF16)
    gelu_bwd_kernel<true><<<grid_for((n + 3) / 4), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(dy, z, out, n);
  else
    gelu_bwd_kernel<false><<<grid_for((n + 3) / 4), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(dy, z, out, n);
  VITB_LAUNCH_CHECK("gelu_bwd_kernel");
  return VITB_OK;
}

int vitb_dropout_fwd(const void* x, int x_dtype, const float* residual, void* y, uint8_t* mask, int64_t n, float p,
                     uint64_t seed, uint64_t* state, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(x && y && mask && state && n > 0, VITB_ERR_BAD_ARG, "dropout_fwd: bad args");
  VITB_REQUIRE(p >= 0.f && p < 1.f, VITB_ERR_BAD_ARG, "dropout_fwd: p=%f must be in [0, 1)", (double)p);
  VITB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & 15u) == 0 &&
               (reinterpret_cast<uintptr_t>(mask) & 3u) == 0 && (reinterpret_cast<uintptr_t>(state) & 7u) == 0,
               VITB_ERR_BAD_ARG, "dropout_fwd: x / y / residual must be 16-byte aligned, mask 4-byte, state 8-byte");
  const double keep = 1.0 - (double)p;
  const double kb = keep * 4294967296.0;
  const uint32_t keep_below = kb >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(kb);
  const float scale = static_cast<float>(1.0 / keep);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = grid_for((n + 3) / 4);
  unsigned long long* sp = reinterpret_cast<unsigned long long*>(state);
  if (x_dtype == VITB_BF16) {
    if (residual) dropout_fwd_kernel<true, true><<<grid, kThreads, 0, stream>>>(x, residual, y, mask, n, keep_below, scale, seed, sp);
    else dropout_fwd_kernel<true, false><<<grid, kThreads, 0, stream>>>(x, residual, y, mask, n, keep_below, scale, seed, sp);
  } else {
    if (residual) dropout_fwd_kernel<false, true><<<grid, kThreads, 0, stream>>>(x, residual, y, mask, n, keep_below, scale, seed, sp);
    else dropout_fwd_kernel<false, false><<<grid, kThreads, 0, stream>>>(x, residual, y, mask, n, keep_below, scale, seed, sp);
  }
  VITB_LAUNCH_CHECK("dropout_fwd_kernel");
  return VITB_OK;
}

int vitb_dropout_bwd(const void* dy, int dy_dtype, const uint8_t* mask, void* dx, int dx_dtype, int64_t n, float p,
                     void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(dy && mask && dx && n > 0 && p >= 0.f && p < 1.f, VITB_ERR_BAD_ARG, "dropout_bwd: bad args");
  const float scale = static_cast<float>(1.0 / (1.0 - (double)p));
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = grid_for(n);
  const bool ib = dy_dtype == VITB_BF16, ob = dx_dtype == VITB_BF16;
  if (ib && ob) dropout_bwd_kernel<true, true><<<grid, kThreads, 0, stream>>>(dy, mask, dx, n, scale);
  else if (ib) dropout_bwd_kernel<true, false><<<grid, kThreads, 0, stream>>>(dy, mask, dx, n, scale);
  else if (ob) dropout_bwd_kernel<false, true><<<grid, kThreads, 0, stream>>>(dy, mask, dx, n, scale);
  else dropout_bwd_kernel<false, false><<<grid, kThreads, 0, stream>>>(dy, mask, dx, n, scale);
  VITB_LAUNCH_CHECK("dropout_bwd_kernel");
  return VITB_OK;
}

int vitb_cast_split(const float* x, int64_t n, void* hi, void* lo, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(x && hi && n > 0, VITB_ERR_BAD_ARG, "cast_split: bad args");
  cast_split_kernel<<<grid_for((n + 3) / 4), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      x, n, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo));
  VITB_LAUNCH_CHECK("cast_split_kernel");
  return VITB_OK;
}

int vitb_im2col(const float* img, int B, int C, int H, int W, int P, int ldk, void* hi, void* lo,
                void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (B == 0) return VITB_OK;
  VITB_REQUIRE(img && hi && B > 0 && C > 0 && P > 0 && H >= P && W >= P, VITB_ERR_BAD_ARG, "im2col: bad args");
  VITB_REQUIRE(ldk % 8 == 0 && ldk >= C * P * P, VITB_ERR_UNSUPPORTED_SHAPE,
               "im2col: ldk=%d must be a multiple of 8 and >= %d", ldk, C * P * P);
  const int gh = H / P, gw = W / P;
  const long long total = static_cast<long long>(B) * gh * gw * (ldk / 8);
  im2col_kernel<<<grid_for(total, 16), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      img, B, C, H, W, P, gh, gw, ldk, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo));
  VITB_LAUNCH_CHECK("im2col_kernel");
  return VITB_OK;
}

int vitb_cls_rows(float* x, int B, int N, int D, const float* cls, const float* pos, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (B == 0) return VITB_OK;
  VITB_REQUIRE(x && cls && B > 0 && N > 0 && D > 0, VITB_ERR_BAD_ARG, "cls_rows: bad args");
  cls_rows_kernel<<<grid_for(static_cast<long long>(B) * D), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      x, B, N, D, cls, pos);
  VITB_LAUNCH_CHECK("cls_rows_kernel");
  return VITB_OK;
}

int vitb_embed_bwd(const float* dx, int B, int N, int D, float* dpos, float* dcls, float* dbias,
                   void* dpatch_hi, void* dpatch_lo, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (B == 0) return VITB_OK;
  VITB_REQUIRE(dx && B > 0 && N > 0 && D > 0 && D % 4 == 0, VITB_ERR_BAD_ARG, "embed_bwd: bad args");
  embed_bwd_kernel<<<grid_for(static_cast<long long>(N) * (D / 4)), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      dx, B, N, D, dpos, dcls, dbias, reinterpret_cast<__nv_bfloat16*>(dpatch_hi),
      reinterpret_cast<__nv_bfloat16*>(dpatch_lo));
  VITB_LAUNCH_CHECK("embed_bwd_kernel");
  return VITB_OK;
}

int vitb_colsum(const void* x, int x_dtype, int rows, int cols, int64_t ld, float* out, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (rows == 0 || cols == 0) return VITB_OK;
  VITB_REQUIRE(x && out && rows > 0 && cols > 0 && cols % 4 == 0 && ld % 4 == 0, VITB_ERR_UNSUPPORTED_SHAPE,
               "colsum: rows=%d cols=%d ld=%lld (cols, ld must be multiples of 4)", rows, cols, (long long)ld);
  colsum_dispatch(x, x_dtype, rows, cols, ld, out, out, out, cols, reinterpret_cast<cudaStream_t>(stream_));
  VITB_LAUNCH_CHECK("colsum_kernel");
  return VITB_OK;
}

int vitb_colsum3(const void* x, int x_dtype, int rows, int seg_cols, int64_t ld, float* out0, float* out1,
                 float* out2, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (rows == 0 || seg_cols == 0) return VITB_OK;
  const int cols = 3 * seg_cols;
  VITB_REQUIRE(x && out0 && out1 && out2 && rows > 0 && seg_cols % 4 == 0 && ld % 4 == 0, VITB_ERR_UNSUPPORTED_SHAPE,
               "colsum3: rows=%d seg_cols=%d ld=%lld", rows, seg_cols, (long long)ld);
  colsum_dispatch(x, x_dtype, rows, cols, ld, out0, out1, out2, seg_cols, reinterpret_cast<cudaStream_t>(stream_));
  VITB_LAUNCH_CHECK("colsum_kernel");
  return VITB_OK;
}

int vitb_cross_entropy(const float* logits, const int64_t* labels, int B, int C, float* loss,
                       float* dlogits, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(logits && labels && loss && B > 0 && C > 0, VITB_ERR_BAD_ARG, "cross_entropy: bad args");
  ce_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      logits, reinterpret_cast<const long long*>(labels), B, C, loss, dlogits);
  VITB_LAUNCH_CHECK("ce_kernel");
  return VITB_OK;
}

int vitb_sgd_momentum(float* p, const float* g, float* m, int64_t n, float lr, const float* hyper_dev, float momentum,
                      float dampening, float weight_decay, int nesterov, int first_step, void* shadow_hi,
                      void* shadow_lo, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(p && g && n > 0 && n % 4 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "sgd: n=%lld must be a multiple of 4",
               (long long)n);
  sgd_kernel<<<grid_for(n / 4), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      p, g, m, n, lr, hyper_dev, momentum, dampening, weight_decay, nesterov, first_step,
      reinterpret_cast<__nv_bfloat16*>(shadow_hi), reinterpret_cast<__nv_bfloat16*>(shadow_lo));
  VITB_LAUNCH_CHECK("sgd_kernel");
  return VITB_OK;
}

int vitb_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, const float* hyper_dev, float beta1,
               float beta2, float eps, float weight_decay, int step, const int* step_dev, const float* grad_scale_dev,
               void* shadow_hi, void* shadow_lo, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || step_dev), VITB_ERR_BAD_ARG, "adamw: bad args");
  adamw_kernel<<<grid_for(n), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      p, g, m, v, n, lr, hyper_dev, beta1, beta2, eps, weight_decay, step, step_dev, grad_scale_dev,
      reinterpret_cast<__nv_bfloat16*>(shadow_hi), reinterpret_cast<__nv_bfloat16*>(shadow_lo));
  VITB_LAUNCH_CHECK("adamw_kernel");
  return VITB_OK;
}

int vitb_adamw_segments(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper_dev, float eps,
                        const float* grad_scale_dev, void* shadow_hi, void* shadow_lo, const int64_t* seg_end,
                        const int32_t* seg_flag, const int32_t* flags, float* seg_step, int nseg, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(p && g && m && v && n > 0 && hyper_dev && seg_end && seg_flag && seg_step && nseg >= 1, VITB_ERR_BAD_ARG,
               "adamw_segments: bad args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  adamw_advance_kernel<<<(nseg + 255) / 256, 256, 0, stream>>>(seg_flag, flags, seg_step, nseg);
  VITB_LAUNCH_CHECK("adamw_advance_kernel");
  adamw_seg_kernel<<<grid_for(n), kThreads, 0, stream>>>(
      p, g, m, v, n, hyper_dev, eps, grad_scale_dev, reinterpret_cast<__nv_bfloat16*>(shadow_hi),
      reinterpret_cast<__nv_bfloat16*>(shadow_lo), reinterpret_cast<const long long*>(seg_end), seg_flag, flags, seg_step, nseg);
  VITB_LAUNCH_CHECK("adamw_seg_kernel");
  return VITB_OK;
}

int vitb_sumsq(const float* x, int64_t n, float* out, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (n == 0) return VITB_OK;
  VITB_REQUIRE(x && out && n > 0, VITB_ERR_BAD_ARG, "sumsq: bad args");
  sumsq_kernel<<<grid_for(n, 4), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(x, n, out);
  VITB_LAUNCH_CHECK("sumsq_kernel");
  return VITB_OK;
}

int vitb_clip_coef(const float* sumsq, float max_norm, float* coef, float* norm_out, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(sumsq && coef, VITB_ERR_BAD_ARG, "clip_coef: bad args");
  clip_coef_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(sumsq, max_norm, coef, norm_out);
  VITB_LAUNCH_CHECK("clip_coef_kernel");
  return VITB_OK;
}

}  // extern "C"
