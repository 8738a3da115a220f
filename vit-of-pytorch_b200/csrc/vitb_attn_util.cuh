// vitb_attn_util.cuh — helpers shared by the tcgen05 attention kernels (vitb_attention_tc.cu: one CTA per tile;
// vitb_attention_ws.cu: persistent warp-specialised kernels): 128B-swizzle addressing of [128 x 64] bf16 tiles,
// TMEM chunk loads, the bf16 staging of an output row for a TMA store.
#pragma once

#include "vitb_common.cuh"

namespace vitb {
namespace attn {

constexpr int DH = 64;
constexpr int kChunkBytes = 128 * 128;  // one 64-key chunk of a [128 x keys] bf16 operand (also one [128 x 64] tile)
constexpr int kAttnThreads = 256;

// byte offset of the 16-byte unit holding keys [8u, 8u+8) of row r inside one 128B-swizzled chunk
__device__ __forceinline__ uint32_t swz_unit(int r, int u) {
  return static_cast<uint32_t>(r * 128 + ((u ^ (r & 7)) << 4));
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ================================================================================================
// Work split inside a CTA (both kernels): 256 threads = 8 warps.  Warp w owns TMEM lanes / query rows
// 32*(w&3) .. +31 (the hardware's lane-quadrant rule) and the column half (w>>2): two threads share a
// row, each handling half of the key columns; row max / row sum are exchanged through shared memory.
// ================================================================================================
struct ColRange { int c_begin, c_end; };   // in units of 32-column chunks (the last chunk may hold 16)
__device__ __forceinline__ ColRange my_chunks(int NK, int half) {
  const int n = (NK + 31) >> 5;
  const int mid = (n + 1) >> 1;
  ColRange r;
  r.c_begin = half ? mid : 0;
  r.c_end = half ? n : mid;
  return r;
}

// loads 32 (or the final 16) fp32 columns of this thread's TMEM lane; missing columns read as `fill`
__device__ __forceinline__ void ld_chunk(uint32_t taddr, int c0, int NK, uint32_t fill, uint32_t (&v)[32]) {
  if (c0 + 32 <= NK) {
    tmem_ld_32x32b_x32(taddr + c0, v);
    tmem_ld_wait();
  } else {
    tmem_ld_32x32b_x16(taddr + c0, reinterpret_cast<uint32_t(&)[16]>(v));   // lands in v[0..16)
    tmem_ld_wait();                                                          // registers are valid only now
#pragma unroll
    for (int j = 16; j < 32; ++j) v[j] = fill;
  }
}

// issue (without waiting) the TMEM load of a 32-column chunk (the final chunk may hold only 16 columns)
__device__ __forceinline__ void issue_chunk(uint32_t taddr, int c0, int NK, uint32_t (&v)[32]) {
  if (c0 + 32 <= NK) {
    tmem_ld_32x32b_x32(taddr + c0, v);
  } else {
    tmem_ld_32x32b_x16(taddr + c0, reinterpret_cast<uint32_t(&)[16]>(v));   // v[16..32) stay unused (masked)
  }
}
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], int c0, int NK, int N) {
  float m = -INFINITY;
  if (c0 + 32 <= N) {   // interior chunk: every column is a valid key
#pragma unroll
    for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) m = fmaxf(m, __uint_as_float(v[j]));
  }
  return m;
}


// this thread's 32 of the 64 head-dim columns of row r -> bf16 -> a 128B-swizzled [128 rows x 64] staging tile that a
// TMA store then writes out (whole 128-byte rows, rows past the token count clipped by the tensor map)
__device__ __forceinline__ void stage_row32_bf16(uint32_t tile, int r, int half, const uint32_t (&v)[32], float scale) {
#pragma unroll
  for (int u = 0; u < 4; ++u)
    st_shared_v4(tile + swz_unit(r, half * 4 + u),
                 pack_bf16x2(__uint_as_float(v[8 * u + 0]) * scale, __uint_as_float(v[8 * u + 1]) * scale),
                 pack_bf16x2(__uint_as_float(v[8 * u + 2]) * scale, __uint_as_float(v[8 * u + 3]) * scale),
                 pack_bf16x2(__uint_as_float(v[8 * u + 4]) * scale, __uint_as_float(v[8 * u + 5]) * scale),
                 pack_bf16x2(__uint_as_float(v[8 * u + 6]) * scale, __uint_as_float(v[8 * u + 7]) * scale));
}

__device__ __forceinline__ float dot8_bf16(const uint4& x, const uint4& y) {
  return bf16_lo(x.x) * bf16_lo(y.x) + bf16_hi(x.x) * bf16_hi(y.x) + bf16_lo(x.y) * bf16_lo(y.y) +
         bf16_hi(x.y) * bf16_hi(y.y) + bf16_lo(x.z) * bf16_lo(y.z) + bf16_hi(x.z) * bf16_hi(y.z) +
         bf16_lo(x.w) * bf16_lo(y.w) + bf16_hi(x.w) * bf16_hi(y.w);
}

}  // namespace attn
}  // namespace vitb
