// vitb_image_prep_core.h — the per-CTA body of the input-transform kernel (vitb_image_prep.cu).
//
// The body is written once, as two phases separated by a CTA barrier, over an explicit
// (block, thread, thread-count, scratch) tuple.  The sm_100a kernel calls it with blockIdx / threadIdx and
// shared memory; tests/host_harness/image_prep_host.cpp compiles the SAME body with g++ and walks the blocks
// and threads in a loop, so the indexing and the integer arithmetic are checked on a machine without a GPU.
// That harness is test scaffolding: nothing in libvitb200.so runs this code on the host.
//
// Arithmetic (bit-exact restatement target: Pillow src/libImaging/Resample.c, 8 bits per channel):
//   acc = 1 << 21;  acc += pixel * weight (weights in 22-bit fixed point);  out = clamp(acc >> 22, 0, 255)
// horizontal pass first, its result rounded to a byte, then the vertical pass; a pass whose axis keeps its
// size is skipped (tables == nullptr).  The normalised float is a 256-entry table lookup per channel.
#pragma once

#include <cuda_bf16.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define VITB_HD __host__ __device__ __forceinline__
#else
#define VITB_HD inline
#endif

namespace vitb_prep {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct Args {
  const uint8_t* src;      // [B, H, W, C] uint8 (what the dataset decodes to)
  int B, H, W, C;
  int out_h, out_w;
  const int32_t* xb;       // [out_w, 2] (first source column, taps) or nullptr when out_w == W
  const int32_t* xc;       // [out_w, xk] fixed-point weights
  int xk;
  const int32_t* yb;       // [out_h, 2] or nullptr when out_h == H
  const int32_t* yc;       // [out_h, yk]
  int yk;
  const uint8_t* flip;     // [B] non-zero = mirror the width axis of the resized image, or nullptr
  const float* lut;        // [C, 256] normalised value of each byte
  float* out_img;          // [B, C, out_h, out_w] fp32 or nullptr
  int vec4_img;            // out_img rows are 16-byte aligned (out_w % 4 == 0 and aligned base)
  int P, ldk, gh, gw;      // patch geometry of the GEMM-operand output (rows (b,py,px), k = (c,ph,pw))
  __nv_bfloat16* cols_hi;  // [B*gh*gw, ldk] bf16 or nullptr
  __nv_bfloat16* cols_lo;  // bf16(v - bf16(v)) (fp32-parity operand) or nullptr
  uint8_t* out_u8;         // [B, out_h, out_w, C] resized bytes after the flip, or nullptr
  int band_rows;           // output rows per CTA
  int rows_cap;            // source rows the scratch can hold per band
};

VITB_HD int clip8(int acc) {
  const int v = acc >> kPrecisionBits;   // arithmetic shift
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// scratch layout: uint8 tmp[rows_cap][C][out_w], padded to 16 bytes, then float lut[C][256]
VITB_HD size_t tmp_bytes(int rows_cap, int C, int out_w) {
  return (static_cast<size_t>(rows_cap) * C * out_w + 15) & ~static_cast<size_t>(15);
}
VITB_HD size_t scratch_bytes(int rows_cap, int C, int out_w) {
  return tmp_bytes(rows_cap, C, out_w) + static_cast<size_t>(C) * 256 * sizeof(float);
}

struct Band {
  int b, oy0, oy1, r0, rows;
};

// Identical on every thread of the CTA: the band of output rows and the source rows it reads.
VITB_HD Band band_of(const Args& a, int block) {
  const int bands = (a.out_h + a.band_rows - 1) / a.band_rows;
  Band bd;
  bd.b = block / bands;
  bd.oy0 = (block - bd.b * bands) * a.band_rows;
  bd.oy1 = bd.oy0 + a.band_rows < a.out_h ? bd.oy0 + a.band_rows : a.out_h;
  int r0, r1;
  if (a.yb) {
    r0 = a.H;
    r1 = 0;
    for (int oy = bd.oy0; oy < bd.oy1; ++oy) {
      const int f = a.yb[2 * oy], n = a.yb[2 * oy + 1];
      r0 = f < r0 ? f : r0;
      r1 = f + n > r1 ? f + n : r1;
    }
  } else {
    r0 = bd.oy0;
    r1 = bd.oy1;
  }
  r0 = r0 < 0 ? 0 : r0;
  r1 = r1 > a.H ? a.H : r1;
  int rows = r1 - r0;
  rows = rows > a.rows_cap ? a.rows_cap : rows;   // memory safety against foreign tables
  bd.r0 = r0;
  bd.rows = rows < 0 ? 0 : rows;
  return bd;
}

// Phase 1: horizontal pass of the band's source rows into scratch (planar per row: [r][c][x]); table copy.
// A warp owns one (source row, channel) line at a time and its lanes walk the output columns, so the only integer
// division is one per line (nthreads is a multiple of 32).
VITB_HD void phase1(const Args& a, int block, int tid, int nthreads, uint8_t* scratch) {
  const Band bd = band_of(a, block);
  float* lut_s = reinterpret_cast<float*>(scratch + tmp_bytes(a.rows_cap, a.C, a.out_w));
  for (int i = tid; i < a.C * 256; i += nthreads) lut_s[i] = a.lut[i];
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int lines = bd.rows * a.C;
  const uint8_t* img = a.src + static_cast<size_t>(bd.b) * a.H * a.W * a.C;
  for (int ln = warp; ln < lines; ln += nwarps) {
    const int r = ln / a.C;
    const int c = ln - r * a.C;
    const uint8_t* row = img + static_cast<size_t>(bd.r0 + r) * a.W * a.C + c;
    uint8_t* dst = scratch + static_cast<size_t>(ln) * a.out_w;
    for (int sx = lane; sx < a.out_w; sx += 32) {
      int v;
      if (a.xb) {
        const int first = a.xb[2 * sx], n = a.xb[2 * sx + 1];
        const int32_t* w = a.xc + static_cast<size_t>(sx) * a.xk;
        int acc = 1 << (kPrecisionBits - 1);
        for (int t = 0; t < n && t < a.xk; ++t) {
          int xi = first + t;
          xi = xi < 0 ? 0 : (xi >= a.W ? a.W - 1 : xi);
          acc += static_cast<int>(row[static_cast<size_t>(xi) * a.C]) * w[t];
        }
        v = clip8(acc);
      } else {
        v = row[static_cast<size_t>(sx) * a.C];
      }
      dst[sx] = static_cast<uint8_t>(v);
    }
  }
}

VITB_HD float bf16_round_f(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// Phase 2: vertical pass, flip, table lookup, and the three optional outputs.  A warp owns one (output row, channel)
// line at a time — its window, weights and patch coordinates are computed once per line — and each lane produces four
// consecutive output columns per step: one 16-byte store into the image (512 contiguous bytes per warp and step), one
// 8-byte store into the patch operand.
VITB_HD void phase2(const Args& a, int block, int tid, int nthreads, const uint8_t* scratch) {
  const Band bd = band_of(a, block);
  const float* lut_s = reinterpret_cast<const float*>(scratch + tmp_bytes(a.rows_cap, a.C, a.out_w));
  const int nq = (a.out_w + 3) >> 2;
  const bool flip = a.flip != nullptr && a.flip[bd.b] != 0;
  const int per_row = a.C * a.out_w;
  const bool vec_cols = (a.P & 3) == 0 && (a.ldk & 3) == 0;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int lines = (bd.oy1 - bd.oy0) * a.C;
  for (int ln = warp; ln < lines; ln += nwarps) {
    const int c = ln % a.C;
    const int oy = bd.oy0 + ln / a.C;
    int yfirst = oy, yn = 1;
    const int32_t* w = nullptr;
    if (a.yb) {
      yfirst = a.yb[2 * oy];
      yn = a.yb[2 * oy + 1];
      yn = yn > a.yk ? a.yk : yn;
      w = a.yc + static_cast<size_t>(oy) * a.yk;
    }
    const float* lut_c = lut_s + c * 256;
    const uint8_t* col_base = scratch + c * a.out_w;       // + row * per_row + sx
    float* img_row = a.out_img ? a.out_img + ((static_cast<size_t>(bd.b) * a.C + c) * a.out_h + oy) * a.out_w : nullptr;
    uint8_t* u8_row = a.out_u8 ? a.out_u8 + (static_cast<size_t>(bd.b) * a.out_h + oy) * a.out_w * a.C + c : nullptr;
    const bool cols_line = a.cols_hi != nullptr && oy < a.gh * a.P;
    size_t cols_row_base = 0;
    int kbase = 0;
    if (cols_line) {
      const int py = oy / a.P, ph = oy - py * a.P;
      cols_row_base = (static_cast<size_t>(bd.b) * a.gh + py) * a.gw;
      kbase = (c * a.P + ph) * a.P;
    }
    for (int q = lane; q < nq; q += 32) {
      int v8[4];
      float val[4];
      const int ox0 = q << 2;
      const int cnt = a.out_w - ox0 < 4 ? a.out_w - ox0 : 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v8[j] = 0;
        val[j] = 0.f;
        if (j < cnt) {
          const int ox = ox0 + j;
          const int sx = flip ? a.out_w - 1 - ox : ox;
          int v;
          if (w) {
            int acc = 1 << (kPrecisionBits - 1);
            for (int t = 0; t < yn; ++t) {
              int rr = yfirst + t - bd.r0;
              rr = rr < 0 ? 0 : (rr >= bd.rows ? bd.rows - 1 : rr);
              acc += static_cast<int>(col_base[rr * per_row + sx]) * w[t];
            }
            v = clip8(acc);
          } else {
            v = col_base[(oy - bd.r0) * per_row + sx];
          }
          v8[j] = v;
          val[j] = lut_c[v];
        }
      }
      if (img_row) {
        float* dst = img_row + ox0;
        if (cnt == 4 && a.vec4_img) {
          *reinterpret_cast<float4*>(dst) = make_float4(val[0], val[1], val[2], val[3]);
        } else {
          for (int j = 0; j < cnt; ++j) dst[j] = val[j];
        }
      }
      if (u8_row) {
        uint8_t* dst = u8_row + static_cast<size_t>(ox0) * a.C;
        for (int j = 0; j < cnt; ++j) dst[static_cast<size_t>(j) * a.C] = static_cast<uint8_t>(v8[j]);
      }
      if (cols_line) {
        if (vec_cols && cnt == 4 && ox0 + 3 < a.gw * a.P) {
          const int px = ox0 / a.P, pw = ox0 - px * a.P;   // P % 4 == 0: the four columns share a patch
          const size_t off = (cols_row_base + px) * a.ldk + kbase + pw;
          __nv_bfloat162 h01 = __floats2bfloat162_rn(val[0], val[1]);
          __nv_bfloat162 h23 = __floats2bfloat162_rn(val[2], val[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h01);
          pk.y = *reinterpret_cast<uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(a.cols_hi + off) = pk;
          if (a.cols_lo) {
            __nv_bfloat162 l01 = __floats2bfloat162_rn(val[0] - bf16_round_f(val[0]), val[1] - bf16_round_f(val[1]));
            __nv_bfloat162 l23 = __floats2bfloat162_rn(val[2] - bf16_round_f(val[2]), val[3] - bf16_round_f(val[3]));
            pk.x = *reinterpret_cast<uint32_t*>(&l01);
            pk.y = *reinterpret_cast<uint32_t*>(&l23);
            *reinterpret_cast<uint2*>(a.cols_lo + off) = pk;
          }
        } else {
          for (int j = 0; j < cnt; ++j) {
            const int ox = ox0 + j;
            if (ox >= a.gw * a.P) break;
            const int px = ox / a.P, pw = ox - px * a.P;
            const size_t off = (cols_row_base + px) * a.ldk + kbase + pw;
            a.cols_hi[off] = __float2bfloat16_rn(val[j]);
            if (a.cols_lo) a.cols_lo[off] = __float2bfloat16_rn(val[j] - bf16_round_f(val[j]));
          }
        }
      }
    }
  }
}

// Host side: band height and scratch rows.  A band of `band` consecutive output rows reads source rows
// [first(oy0), first(oy1-1) + taps): with first = int(center - support + .5) and end = int(center + support + .5)
// that span is at most (band-1)*scale + 2*support + 1 rows.
inline size_t choose_band(int H, int out_h, bool vertical_pass, int C, int out_w, int* band_rows, int* rows_cap) {
  const double scale = static_cast<double>(H) / static_cast<double>(out_h);
  const double support = scale < 1.0 ? 1.0 : scale;
  int band = 32, cap = 0;
  size_t bytes = 0;
  for (;; band >>= 1) {
    cap = vertical_pass ? static_cast<int>(ceil((band - 1) * scale + 2.0 * support)) + 2 : band;
    if (cap > H) cap = H;
    bytes = scratch_bytes(cap, C, out_w);
    if (bytes <= 64 * 1024 || band == 1) break;
  }
  *band_rows = band;
  *rows_cap = cap;
  return bytes;
}

}  // namespace vitb_prep
