// vitb_compact.cu — device-side row compaction for Res-ViT's inference-time token skipping.
//
// In eval mode a dynamic TransformerBlock keeps the attention / MLP result only for the tokens its router marked active
// and passes the others through (res-vit/model.py:503-524: `student_out = mask * output + (~mask) * x`).  The reference
// computes the output projection's input per image with boolean indexing (one host sync per image) and the MLP on all
// rows.  Here the active rows are compacted ON THE DEVICE — no count ever travels to the host: the row list and its
// length stay in device memory, the GEMMs behind it take their row count from there (vitb_gemm_params.m_dev), and the
// results are scattered back over a copy of x:
//     vitb_compact_rows : rows[0 .. *count) = { t : member(index[t]) }            (order unspecified; *count += ...)
//     vitb_gather_rows  : dst[i, :] = src[rows[i], :]      for i < *count
//     vitb_scatter_rows : dst[rows[i], :] = src[i, :]      for i < *count
// All of them are HBM-bound copies with 16-byte accesses; the row order inside the list does not matter to anything
// that follows (every consumer is row-wise), so the list is built with one atomic per warp instead of a scan.
#include "../../include/vitb200.h"
#include "vitb_common.cuh"

namespace {
using namespace vitb;

constexpr int kThreads = 256;

inline int grid_for(long long items, int per_sm = 8) {
  long long blocks = (items + kThreads - 1) / kThreads;
  const long long cap = static_cast<long long>(vitb_num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

__global__ void __launch_bounds__(kThreads)
compact_rows_kernel(const float* __restrict__ index, unsigned mask, int T, int* __restrict__ rows, int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  const int t_end = (T + 31) & ~31;   // whole warps iterate together (ballot)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < t_end; t += stride) {
    bool sel = false;
    if (t < T) {
      const int idx = static_cast<int>(index[t]);
      sel = (idx >= 0 && idx < 32) ? ((mask >> idx) & 1u) != 0 : false;
    }
    const unsigned b = __ballot_sync(0xffffffffu, sel);
    int base = 0;
    if (lane == 0 && b != 0) base = atomicAdd(count, __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (sel) rows[base + __popc(b & ((1u << lane) - 1u))] = t;
  }
}

// GATHER: dst[i] = src[rows[i]];  !GATHER: dst[rows[i]] = src[i].  16-byte vectors, cv of them per row.
template <bool GATHER>
__global__ void __launch_bounds__(kThreads)
move_rows_kernel(const uint4* __restrict__ src, long long src_ld16, uint4* __restrict__ dst, long long dst_ld16,
                 const int* __restrict__ rows, const int* __restrict__ count, int cv) {
  const long long total = static_cast<long long>(*count) * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cv), c = static_cast<int>(i - static_cast<long long>(r) * cv);
    const int t = rows[r];
    if (GATHER) dst[static_cast<long long>(r) * dst_ld16 + c] = src[static_cast<long long>(t) * src_ld16 + c];
    else dst[static_cast<long long>(t) * dst_ld16 + c] = src[static_cast<long long>(r) * src_ld16 + c];
  }
}

int check_move(const void* src, const void* dst, const int* rows, const int* count, int cols, int esize, int64_t src_ld,
               int64_t dst_ld, const char* who) {
  VITB_REQUIRE(src && dst && rows && count && cols > 0, VITB_ERR_BAD_ARG, "%s: bad args", who);
  const int V = 16 / esize;
  VITB_REQUIRE(cols % V == 0 && src_ld % V == 0 && dst_ld % V == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) & 15u) == 0,
               VITB_ERR_UNSUPPORTED_SHAPE, "%s: rows must be 16-byte aligned multiples of 16 bytes", who);
  return VITB_OK;
}

}  // namespace

extern "C" {

int vitb_compact_rows(const float* index, uint32_t member_mask, int T, int32_t* rows, int32_t* count, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (T == 0) return VITB_OK;
  VITB_REQUIRE(index && rows && count && T > 0, VITB_ERR_BAD_ARG, "compact_rows: bad args");
  compact_rows_kernel<<<grid_for(T, 4), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(index, member_mask, T, rows, count);
  VITB_LAUNCH_CHECK("compact_rows_kernel");
  return VITB_OK;
}

int vitb_gather_rows(const void* src, int64_t src_ld, int dtype, const int32_t* rows, const int32_t* count, int max_rows,
                     int cols, void* dst, int64_t dst_ld, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (max_rows == 0) return VITB_OK;
  const int es = dtype == VITB_BF16 ? 2 : 4;
  if ((st = check_move(src, dst, rows, count, cols, es, src_ld, dst_ld, "gather_rows")) != VITB_OK) return st;
  const int cv = cols * es / 16;
  move_rows_kernel<true><<<grid_for(static_cast<long long>(max_rows) * cv), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const uint4*>(src), src_ld * es / 16, reinterpret_cast<uint4*>(dst), dst_ld * es / 16, rows, count, cv);
  VITB_LAUNCH_CHECK("gather_rows_kernel");
  return VITB_OK;
}

int vitb_scatter_rows(const void* src, int64_t src_ld, int dtype, const int32_t* rows, const int32_t* count, int max_rows,
                      int cols, void* dst, int64_t dst_ld, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  if (max_rows == 0) return VITB_OK;
  const int es = dtype == VITB_BF16 ? 2 : 4;
  if ((st = check_move(src, dst, rows, count, cols, es, src_ld, dst_ld, "scatter_rows")) != VITB_OK) return st;
  const int cv = cols * es / 16;
  move_rows_kernel<false><<<grid_for(static_cast<long long>(max_rows) * cv), kThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const uint4*>(src), src_ld * es / 16, reinterpret_cast<uint4*>(dst), dst_ld * es / 16, rows, count, cv);
  VITB_LAUNCH_CHECK("scatter_rows_kernel");
  return VITB_OK;
}

}  // extern "C"
