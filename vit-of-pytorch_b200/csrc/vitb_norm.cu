// vitb_norm.cu — LayerNorm forward/backward for the ViT encoder path (HBM-bound kernels).
//
// Replaces nn.LayerNorm(D, eps) at src/model.py:108,114,146 (calls :119,:127,:155) and the wrapper
// at res-vit/model.py:119-130, together with their autograd backward, and folds in the work that
// surrounds them on the path: the bf16 (or bf16 hi/lo split) GEMM operand is emitted directly, the
// residual-branch gradient is added on the fly, and the column sums the neighbouring bias
// gradients need are accumulated in the same pass.
//
// Mapping: one warp per row, the row lives in registers as NV float4 per lane (D = 128*NV), 128-bit
// coalesced global access, warp-shuffle reductions, fp32 statistics (two-pass, like ATen).
// Algorithmic bytes per row (D columns): fwd 4D (x) + 2D (y bf16) + 8; bwd 2D (dy) + 4D (x) +
// 4D (dres) + 4D (dx) + 2D (dx bf16).
#include "../../include/vitb200.h"
#include "vitb_common.cuh"

namespace {
using namespace vitb;

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_bf16(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
}
__device__ __forceinline__ void st4_bf16(__nv_bfloat16* p, float a, float b, float c, float d) {
  uint2 u;
  u.x = pack_bf16x2(a, b);
  u.y = pack_bf16x2(c, d);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

template <int NV, bool X_BF16>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_fwd_kernel(const void* __restrict__ x_, long long x_stride, int rows, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, float* __restrict__ y_f32,
              __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo,
              float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int D = NV * 128;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int nw = gridDim.x * kWarpsPerBlock;
  float4 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i] = ld4(gamma + (i * 32 + lane) * 4);
    b[i] = ld4(beta + (i * 32 + lane) * 4);
  }
  for (int row = gw; row < rows; row += nw) {
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const long long off = static_cast<long long>(row) * x_stride + (i * 32 + lane) * 4;
      if constexpr (X_BF16) v[i] = ld4_bf16(reinterpret_cast<const __nv_bfloat16*>(x_) + off);
      else v[i] = ld4(reinterpret_cast<const float*>(x_) + off);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float y0 = (v[i].x - mean) * rstd * g[i].x + b[i].x;
      const float y1 = (v[i].y - mean) * rstd * g[i].y + b[i].y;
      const float y2 = (v[i].z - mean) * rstd * g[i].z + b[i].z;
      const float y3 = (v[i].w - mean) * rstd * g[i].w + b[i].w;
      const long long off = static_cast<long long>(row) * D + (i * 32 + lane) * 4;
      if (y_f32) *reinterpret_cast<float4*>(y_f32 + off) = make_float4(y0, y1, y2, y3);
      if (y_hi) st4_bf16(y_hi + off, y0, y1, y2, y3);
      if (y_lo)
        st4_bf16(y_lo + off, y0 - bf16_round(y0), y1 - bf16_round(y1), y2 - bf16_round(y2),
                 y3 - bf16_round(y3));
    }
  }
}

// Backward.  One warp per row, the row in registers.  HBM-bound (16 D bytes per row), and what buys bandwidth is
// rows in flight per SM, i.e. warps per SM, i.e. registers per thread.  The first version kept the warp's partial
// dgamma / dbeta / dcolsum sums (12 D bytes per warp) and gamma in registers: 240 registers, 8 warps per SM, 74 % of
// the measured HBM peak (profiles/launches_r01c_summary.txt).  Here the partial sums live in a per-warp slice of
// shared memory (plain 16-byte read-modify-write, no atomics: the slice is private to the warp and every lane owns
// its columns), gamma is re-read from L1 per row (3 KB, always resident) and a bf16 dy stays packed between the two
// phases: 16 warps per SM for D <= 768 (12 for wider rows, whose slices are larger).
template <int NV>
struct LnBwdCfg {
  static constexpr int WARPS = NV <= 6 ? 16 : 12;
  static constexpr size_t SMEM = static_cast<size_t>(WARPS) * 3 * NV * 128 * sizeof(float);
};

template <int NV, bool DY_BF16, bool COLSUM>
__global__ void __launch_bounds__(LnBwdCfg<NV>::WARPS * 32, 1)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, long long x_stride,
              const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* __restrict__ gamma, int rows, const float* __restrict__ dres,
              long long dres_stride, int dres_every, float* __restrict__ dx_f32, long long dx_stride,
              __nv_bfloat16* __restrict__ dx_hi, __nv_bfloat16* __restrict__ dx_lo,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum) {
  constexpr int D = NV * 128;
  constexpr int WARPS = LnBwdCfg<NV>::WARPS;
  extern __shared__ __align__(16) float s_acc[];   // [WARPS][3][D]: dgamma | dbeta | dcolsum partial sums per warp
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* acc = s_acc + static_cast<size_t>(warp) * 3 * D;
#pragma unroll
  for (int i = 0; i < 3 * NV; ++i)
    *reinterpret_cast<float4*>(acc + (i * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  pdl_wait();
  const int gw = blockIdx.x * WARPS + warp;
  const int nw = gridDim.x * WARPS;
  // dy of the row between the two phases: raw bf16x4 (two registers) or fp32x4
  struct DyKeep { uint32_t a, b, c, d; };
  auto dy_value = [](const DyKeep& k) {
    if constexpr (DY_BF16) return make_float4(bf16_lo(k.a), bf16_hi(k.a), bf16_lo(k.b), bf16_hi(k.b));
    else return make_float4(__uint_as_float(k.a), __uint_as_float(k.b), __uint_as_float(k.c), __uint_as_float(k.d));
  };
  auto add4 = [](float* p, float a, float b, float c, float d) {
    float4 v = *reinterpret_cast<float4*>(p);
    v.x += a; v.y += b; v.z += c; v.w += d;
    *reinterpret_cast<float4*>(p) = v;
  };
  for (int row = gw; row < rows; row += nw) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV];
    DyKeep keep[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 xv = ld4(x + static_cast<long long>(row) * x_stride + c);
      if constexpr (DY_BF16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + static_cast<long long>(row) * D + c);
        keep[i].a = u.x; keep[i].b = u.y; keep[i].c = 0u; keep[i].d = 0u;
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dy_) + static_cast<long long>(row) * D + c);
        keep[i].a = u.x; keep[i].b = u.y; keep[i].c = u.z; keep[i].d = u.w;
      }
      xv.x = (xv.x - mu) * rs; xv.y = (xv.y - mu) * rs; xv.z = (xv.z - mu) * rs; xv.w = (xv.w - mu) * rs;
      xh[i] = xv;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 dyv = dy_value(keep[i]);
      const float4 xv = xh[i];
      add4(acc + c, dyv.x * xv.x, dyv.y * xv.y, dyv.z * xv.z, dyv.w * xv.w);        // dgamma
      add4(acc + D + c, dyv.x, dyv.y, dyv.z, dyv.w);                                 // dbeta
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      dyv.x *= gm.x; dyv.y *= gm.y; dyv.z *= gm.z; dyv.w *= gm.w;
      s1 += (dyv.x + dyv.y) + (dyv.z + dyv.w);
      s2 += (dyv.x * xv.x + dyv.y * xv.y) + (dyv.z * xv.z + dyv.w * xv.w);
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
    // residual-branch gradient of this row: every row has one, or (dres_every = n > 1) only rows 0, n, 2n, ...
    const float* dres_row = nullptr;
    if (dres) {
      if (dres_every == 1) dres_row = dres + static_cast<long long>(row) * dres_stride;
      else if (row % dres_every == 0) dres_row = dres + static_cast<long long>(row / dres_every) * dres_stride;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      float4 gy = dy_value(keep[i]);
      gy.x *= gm.x; gy.y *= gm.y; gy.z *= gm.z; gy.w *= gm.w;    // the same products phase 1 summed
      float4 d;
      d.x = rs * (gy.x - s1 - xh[i].x * s2);
      d.y = rs * (gy.y - s1 - xh[i].y * s2);
      d.z = rs * (gy.z - s1 - xh[i].z * s2);
      d.w = rs * (gy.w - s1 - xh[i].w * s2);
      if (dres_row) {
        const float4 r = ld4(dres_row + c);
        d.x += r.x; d.y += r.y; d.z += r.z; d.w += r.w;
      }
      if constexpr (COLSUM) add4(acc + 2 * D + c, d.x, d.y, d.z, d.w);
      if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + static_cast<long long>(row) * dx_stride + c) = d;
      const long long off = static_cast<long long>(row) * D + c;
      if (dx_hi) st4_bf16(dx_hi + off, d.x, d.y, d.z, d.w);
      if (dx_lo)
        st4_bf16(dx_lo + off, d.x - bf16_round(d.x), d.y - bf16_round(d.y), d.z - bf16_round(d.z),
                 d.w - bf16_round(d.w));
    }
  }
  // fold the warps' slices, then one global atomic per column and block
  __syncthreads();
  constexpr int NSUM = COLSUM ? 3 : 2;
  // one scalar reduction per column, consecutive lanes on consecutive columns: a warp's 32 reductions are ONE 128-byte
  // transaction at L2.  (red.global.add.v4.f32 here — four columns per lane — made this kernel 36 % slower, 56 -> 76 us
  // per launch, profiles/launches_r02n.csv: each lane's 16 bytes travel as a transaction of their own.)
  for (int i = threadIdx.x; i < NSUM * D; i += blockDim.x) {
    float t = 0.f;
#pragma unroll 4
    for (int w = 0; w < WARPS; ++w) t += s_acc[static_cast<size_t>(w) * 3 * D + i];
    if (i < D) { if (dgamma) atomicAdd(dgamma + i, t); }
    else if (i < 2 * D) { if (dbeta) atomicAdd(dbeta + i - D, t); }
    else atomicAdd(dcolsum + i - 2 * D, t);
  }
}

#define VITB_NV_SWITCH(nv, ...)                     \
  switch (nv) {                                     \
    case 1: { constexpr int NV = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int NV = 2; __VA_ARGS__; break; }   \
    case 3: { constexpr int NV = 3; __VA_ARGS__; break; }   \
    case 4: { constexpr int NV = 4; __VA_ARGS__; break; }   \
    case 6: { constexpr int NV = 6; __VA_ARGS__; break; }   \
    case 8: { constexpr int NV = 8; __VA_ARGS__; break; }   \
    case 10: { constexpr int NV = 10; __VA_ARGS__; break; } \
    default:                                        \
      vitb_set_error("LayerNorm width %d unsupported (need D/128 in {1,2,3,4,6,8,10})", (nv) * 128); \
      return VITB_ERR_UNSUPPORTED_SHAPE;            \
  }

}  // namespace

extern "C" int vitb_layernorm_fwd(const void* x, int x_dtype, int64_t x_row_stride, int rows, int D,
                                  const float* gamma, const float* beta, float eps, float* y_f32,
                                  void* y_bf16, void* y_bf16_lo, float* mean, float* rstd,
                                  void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (rows == 0) return VITB_OK;
  VITB_REQUIRE(rows > 0 && D > 0 && D % 128 == 0, VITB_ERR_UNSUPPORTED_SHAPE,
               "layernorm_fwd: rows=%d D=%d (D must be a multiple of 128)", rows, D);
  VITB_REQUIRE(x && gamma && beta, VITB_ERR_BAD_ARG, "layernorm_fwd: null input");
  VITB_REQUIRE(x_row_stride % 4 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "layernorm_fwd: row stride %% 4");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int blocks_needed = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int max_blocks = vitb_num_sms() * 4;
  const int grid = blocks_needed < max_blocks ? blocks_needed : max_blocks;
  const int nv = D / 128;
  __nv_bfloat16* yh = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  __nv_bfloat16* yl = reinterpret_cast<__nv_bfloat16*>(y_bf16_lo);
  cudaError_t lerr = cudaSuccess;
  if (x_dtype == VITB_BF16) {
    VITB_NV_SWITCH(nv, (lerr = vitb_launch<kPdlNorm>(ln_fwd_kernel<NV, true>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream,
                                           x, x_row_stride, rows, gamma, beta, eps, y_f32, yh, yl, mean, rstd)));
  } else {
    VITB_NV_SWITCH(nv, (lerr = vitb_launch<kPdlNorm>(ln_fwd_kernel<NV, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, stream,
                                           x, x_row_stride, rows, gamma, beta, eps, y_f32, yh, yl, mean, rstd)));
  }
  VITB_CUDA_CHECK(lerr);
  VITB_LAUNCH_CHECK("ln_fwd_kernel");
  return VITB_OK;
}

static int layernorm_bwd_impl(const void* dy, int dy_dtype, const float* x, int64_t x_row_stride,
                              const float* mean, const float* rstd, const float* gamma, int rows,
                              int D, const float* dres, int64_t dres_row_stride, int dres_every, float* dx_f32,
                              int64_t dx_row_stride, void* dx_bf16, void* dx_bf16_lo,
                              float* dgamma, float* dbeta, float* dcolsum, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (rows == 0) return VITB_OK;
  VITB_REQUIRE(rows > 0 && D > 0 && D % 128 == 0, VITB_ERR_UNSUPPORTED_SHAPE,
               "layernorm_bwd: rows=%d D=%d", rows, D);
  VITB_REQUIRE(dy && x && mean && rstd && gamma, VITB_ERR_BAD_ARG, "layernorm_bwd: null input");
  VITB_REQUIRE(dres_every >= 1, VITB_ERR_BAD_ARG, "layernorm_bwd: dres_every=%d", dres_every);
  VITB_REQUIRE(x_row_stride % 4 == 0 && dres_row_stride % 4 == 0 && dx_row_stride % 4 == 0,
               VITB_ERR_UNSUPPORTED_SHAPE, "layernorm_bwd: row strides must be multiples of 4");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int nv = D / 128;
  const int bw = nv <= 6 ? 16 : 12;                // LnBwdCfg<NV>::WARPS
  const int blocks_needed = (rows + bw - 1) / bw;
  const int max_blocks = vitb_num_sms();          // one block per SM (the warps' slices take 147-184 KB)
  const int grid = blocks_needed < max_blocks ? blocks_needed : max_blocks;
  const size_t smem = static_cast<size_t>(bw) * 3 * D * sizeof(float);
  __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  __nv_bfloat16* dl = reinterpret_cast<__nv_bfloat16*>(dx_bf16_lo);
#define VITB_LNB(DYB, CS)                                                                        \
  VITB_NV_SWITCH(nv, (lerr = cudaFuncSetAttribute(ln_bwd_kernel<NV, DYB, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  static_cast<int>(LnBwdCfg<NV>::SMEM)),                          \
                      lerr = (lerr != cudaSuccess) ? lerr :                                                       \
                             vitb_launch<kPdlNorm>(ln_bwd_kernel<NV, DYB, CS>, dim3(grid), dim3(LnBwdCfg<NV>::WARPS * 32), smem,  \
                                         stream, dy, x, x_row_stride, mean, rstd, gamma, rows, dres,              \
                                         dres_row_stride, dres_every, dx_f32, dx_row_stride, dh, dl, dgamma, dbeta, dcolsum)))
  cudaError_t lerr = cudaSuccess;
  if (dy_dtype == VITB_BF16) {
    if (dcolsum) { VITB_LNB(true, true); } else { VITB_LNB(true, false); }
  } else {
    if (dcolsum) { VITB_LNB(false, true); } else { VITB_LNB(false, false); }
  }
#undef VITB_LNB
  VITB_CUDA_CHECK(lerr);
  VITB_LAUNCH_CHECK("ln_bwd_kernel");
  return VITB_OK;
}

extern "C" int vitb_layernorm_bwd(const void* dy, int dy_dtype, const float* x, int64_t x_row_stride,
                                  const float* mean, const float* rstd, const float* gamma, int rows,
                                  int D, const float* dres, int64_t dres_row_stride, float* dx_f32,
                                  int64_t dx_row_stride, void* dx_bf16, void* dx_bf16_lo,
                                  float* dgamma, float* dbeta, float* dcolsum, void* stream_) {
  return layernorm_bwd_impl(dy, dy_dtype, x, x_row_stride, mean, rstd, gamma, rows, D, dres, dres_row_stride, 1, dx_f32,
                            dx_row_stride, dx_bf16, dx_bf16_lo, dgamma, dbeta, dcolsum, stream_);
}

// Same, with a residual-branch gradient that only every dres_every-th row has: dres row r / dres_every is added to
// row r when r % dres_every == 0 (the class-token-only last encoder block: of the N rows of an image only row 0
// receives a gradient through the residual connection; src/model.py:155,210).
extern "C" int vitb_layernorm_bwd_sparse_res(const void* dy, int dy_dtype, const float* x, int64_t x_row_stride,
                                             const float* mean, const float* rstd, const float* gamma, int rows,
                                             int D, const float* dres, int64_t dres_row_stride, int dres_every,
                                             float* dx_f32, int64_t dx_row_stride, void* dx_bf16, void* dx_bf16_lo,
                                             float* dgamma, float* dbeta, float* dcolsum, void* stream_) {
  return layernorm_bwd_impl(dy, dy_dtype, x, x_row_stride, mean, rstd, gamma, rows, D, dres, dres_row_stride, dres_every,
                            dx_f32, dx_row_stride, dx_bf16, dx_bf16_lo, dgamma, dbeta, dcolsum, stream_);
}
