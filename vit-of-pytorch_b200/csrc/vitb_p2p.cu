// vitb_p2p.cu — the gradient exchange of the data-parallel step as ONE kernel over NVLink peer memory.
//
// Replaces nn.DataParallel's gradient reduction (src/train.py:128-129) / the NCCL all-reduce of the flat fp32 gradient
// buffer (train.GraphedTrainStep).  Every rank holds the buffer in SYMMETRIC memory (same size on every GPU, mapped into
// every process: torch.distributed._symmetric_memory does the allocation and the handle exchange — plumbing), and the
// kernel gets the W peer pointers, the NVSwitch multicast pointer when the fabric offers one, and W signal pads.
//
//   barrier A            every rank's gradients are complete (release / acquire at system scope)
//   reduce + broadcast   rank r owns elements [r n / W, (r + 1) n / W):
//                          multicast:   v = multimem.ld_reduce.add(mc + i)    the SWITCH sums the W replicas (NVLS)
//                                       multimem.st(mc + i, v * scale)        the switch writes all W replicas
//                          peer-to-peer: v = sum_p ld(buf[p] + i);  st(buf[p] + i, v * scale) for every p
//   barrier B            every rank's share has landed everywhere
//
// NVLink traffic per GPU: multicast n (1/W out + (W-1)/W... ~ n) bytes each way against 2 n (W - 1) / W for the
// peer-to-peer form (and for a ring).  The barriers pair block b of every rank with block b of every other rank: slot
// [b * W + src] of a rank's signal pad counts the barrier rounds that rank src's block b has entered; a block waits until
// all W counters of its row have reached its own round number.  Counters only grow, nothing is ever reset, so the kernel
// is replayable from a CUDA graph; the round number lives in the pad as well (slot kRoundBase + b, local).
// All blocks of all ranks must be co-resident: the grid is at most one block per SM and nothing else runs on the stream.
#include <stdlib.h>

#include "../../include/vitb200.h"
#include "vitb_common.cuh"

namespace {
using namespace vitb;

constexpr int kP2pThreads = 512;
constexpr int kMaxWorld = 16;
constexpr int kMaxBlocks = 160;
constexpr int kUnroll = 4;                            // 16-byte vectors in flight per thread: remote latency is microseconds
constexpr int kRoundBase = kMaxBlocks * kMaxWorld;     // pad layout: [kMaxBlocks][kMaxWorld] counters | [kMaxBlocks] rounds

struct P2pArgs {
  float* buf[kMaxWorld];
  unsigned* pad[kMaxWorld];
  float* mc;               // multicast address of the buffer, or null
  long long n;             // fp32 elements, a multiple of 4
  float scale;
  int rank, world;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys(unsigned* p) {
  asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// block b of this rank meets block b of every other rank
__device__ __forceinline__ void cross_gpu_barrier(const P2pArgs& a, unsigned round) {
  __syncthreads();                                   // this block's memory operations are issued ...
  if (threadIdx.x < a.world) {
    __threadfence_system();                          // ... and ordered before the signal, system-wide
    red_release_sys(a.pad[threadIdx.x] + blockIdx.x * kMaxWorld + a.rank);
    const unsigned* mine = a.pad[a.rank] + blockIdx.x * kMaxWorld + threadIdx.x;
    long long spins = 0;
    while (ld_acquire_sys(mine) < round) {
      if (++spins > (1ll << 30)) __trap();           // a lost peer must trap (after ~10 minutes: ranks may be skewed by seconds), never hang the box
      __nanosleep(40);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kP2pThreads, 1)
p2p_allreduce_kernel(const P2pArgs a) {
  unsigned* round_slot = a.pad[a.rank] + kRoundBase + blockIdx.x;
  const unsigned round0 = *reinterpret_cast<volatile unsigned*>(round_slot);    // barrier rounds this block has completed
  cross_gpu_barrier(a, round0 + 1);
  const long long per = ((a.n / 4 + a.world - 1) / a.world) * 4;               // elements owned per rank, 16-byte units
  const long long lo = per * a.rank, hi = min(a.n, lo + per);
  // a block covers kUnroll * blockDim.x consecutive vectors per step; a thread's kUnroll vectors are blockDim.x apart, so
  // every load instruction of a warp is 512 contiguous bytes and kUnroll of them are in flight before the first store
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x * 4 * kUnroll;
  const long long first = lo + (static_cast<long long>(blockIdx.x) * blockDim.x * kUnroll + threadIdx.x) * 4;
  if (a.mc != nullptr) {
    for (long long i = first; i < hi; i += step) {
      float4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = i + static_cast<long long>(u) * blockDim.x * 4;
        if (j < hi)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(a.mc + j) : "memory");
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = i + static_cast<long long>(u) * blockDim.x * 4;
        if (j < hi)
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc + j), "f"(v[u].x * a.scale),
                       "f"(v[u].y * a.scale), "f"(v[u].z * a.scale), "f"(v[u].w * a.scale) : "memory");
      }
    }
  } else {
    for (long long i = first; i < hi; i += step) {
      float4 acc[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < a.world; ++p) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long j = i + static_cast<long long>(u) * blockDim.x * 4;
          if (j < hi) {
            const float4 v = __ldcv(reinterpret_cast<const float4*>(a.buf[p] + j));     // never from a stale L1 line
            acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = i + static_cast<long long>(u) * blockDim.x * 4;
        if (j < hi) {
          const float4 r = make_float4(acc[u].x * a.scale, acc[u].y * a.scale, acc[u].z * a.scale, acc[u].w * a.scale);
          for (int p = 0; p < a.world; ++p) *reinterpret_cast<float4*>(a.buf[p] + j) = r;
        }
      }
    }
  }
  cross_gpu_barrier(a, round0 + 2);
  if (threadIdx.x == 0) *round_slot = round0 + 2;
}
}  // namespace

extern "C" int vitb_p2p_pad_words(void) { return kRoundBase + kMaxBlocks; }

extern "C" int vitb_p2p_allreduce(float* const* bufs, uint32_t* const* pads, float* multicast, int rank, int world, int64_t n,
                                  float scale, void* stream_) {
  int st = vitb_check_device();
  if (st != VITB_OK) return st;
  VITB_REQUIRE(bufs && pads && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, VITB_ERR_BAD_ARG,
               "p2p_allreduce: world=%d rank=%d (at most %d ranks)", world, rank, kMaxWorld);
  VITB_REQUIRE(n > 0 && n % 4 == 0, VITB_ERR_UNSUPPORTED_SHAPE, "p2p_allreduce: n=%lld must be a positive multiple of 4", (long long)n);
  P2pArgs a{};
  for (int p = 0; p < world; ++p) {
    VITB_REQUIRE(bufs[p] && pads[p] && (reinterpret_cast<uintptr_t>(bufs[p]) & 15u) == 0, VITB_ERR_BAD_ARG,
                 "p2p_allreduce: peer %d buffer / pad missing or not 16-byte aligned", p);
    a.buf[p] = bufs[p];
    a.pad[p] = pads[p];
  }
  VITB_REQUIRE((reinterpret_cast<uintptr_t>(multicast) & 15u) == 0, VITB_ERR_BAD_ARG, "p2p_allreduce: multicast pointer alignment");
  a.mc = multicast;
  a.n = n; a.scale = scale; a.rank = rank; a.world = world;
  int grid = vitb_num_sms();                          // one block per SM: all blocks of all ranks must be co-resident
  if (const char* e = getenv("VITB_P2P_BLOCKS")) { const int g = atoi(e); if (g >= 1) grid = g; }   // tuning runs
  if (grid > kMaxBlocks) grid = kMaxBlocks;
  if (grid < 1) grid = 1;
  p2p_allreduce_kernel<<<grid, kP2pThreads, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(a);
  VITB_LAUNCH_CHECK("p2p_allreduce_kernel");
  return VITB_OK;
}
