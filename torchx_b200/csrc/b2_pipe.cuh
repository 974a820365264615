// b2_pipe.cuh — chunk-pipelined, warp-specialised allreduce kernels.
//
// Why: in the single-pass two-shot kernel (b2_kernels.cuh) every CTA runs scatter -> barrier -> reduce -> barrier -> gather
// in lock step, so the NVLink phases, the HBM-only phases and the ~5-10 us barriers (release fence = drain of every
// outstanding remote store, flag round trip) add up: at a 25 MiB DDP bucket and W = 8 only half of the 59 us is wire
// time (profiles/r01_phase_trace_w4.md).  Here the message is cut into K chunks and the CTA into three ROLE GROUPS of
// warps that run the three phases of DIFFERENT chunks at the same time and never synchronise with each other inside the
// CTA - the only dependencies are the cross-rank ones, carried by per-(CTA index, kind, chunk) flag words:
//
//   NVLS  (k_pipe<.., kNvls>; needs the multicast mapping of the arena):
//     A cast   : read my fp32 bucket once, cast+scale, store bf16 into MY stage (local HBM/L2)        -> signal X1[k]
//     B move   : wait X1[k] from all ranks; multimem.ld_reduce my slice (the SWITCH sums the W stages, fp32 accumulate,
//                one rounding) and multimem.st the result back to every rank's stage (in place)       -> signal X2[k]
//     C widen  : wait X2[k] from all ranks; read my stage (all W slices, now reduced), widen, write my bucket
//     NVLink traffic per GPU and direction: (1 + 1/W) * S instead of the 2 (W-1)/W * S of any P2P algorithm.
//
//   P2P   (k_pipe<.., kP2p>; any world, plain peer mappings):
//     A scatter: read my bucket once, cast+scale, STORE slice j into rank j's recv[me]                -> signal X1[k]
//     B reduce : wait X1[k]; sum recv[0..W-1] of my slice in rank order (fp32), round once -> "reduced" -> signal X2[k]
//     C gather : wait X2[k]; LOAD slice j from rank j's "reduced", widen, write my bucket
//
// A role group signals with  bar.sync(group) -> st.release.sys(flag)  by its first W threads and waits with
// ld.acquire.sys(flag) -> bar.sync(group); the release fence therefore only ever drains ONE chunk of one role's stores
// while the other two roles keep the memory system busy.  Chunk (k, b) = vecs [(k*g + b) * cell, +cell) of every slice
// (g = grid size): chunk-major, so the bytes of one pipeline step are contiguous bands of each slice, and CTA b of every
// rank touches exactly the same vec indices - the only cross-rank dependencies are between equal CTA indices, so neither
// a grid-wide sync nor co-residency of a rank's own CTAs is needed.
//
// Staging-buffer reuse across collectives is safe for the same reason as in the single-pass kernels:
// the two stages alternate by the parity of the device-resident op counter, and collective n+2 can only start once every
// CTA of collective n+1 has passed its X2 waits, which peers signal only after their collective n has completed.
#pragma once

#include "b2_dev.cuh"

namespace pl {

enum { kNvls = 0, kP2p = 1 };

// role group sizes (threads): A = cast/scatter, B = move/reduce, C = widen/gather
constexpr int kA = 192, kB = 128, kC = 192;
static_assert(kA + kB + kC == kThreads, "role groups must tile the CTA");
static_assert(kA % 32 == 0 && kB % 32 == 0 && kC % 32 == 0, "role groups are whole warps");

__device__ __forceinline__ void group_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ uint32_t* flag_slot(uint8_t* arena, const CommDev& c, int kind, int k) {
  const size_t idx = (static_cast<size_t>(blockIdx.x) * kPipeKinds + kind) * kMaxChunks + k;
  return reinterpret_cast<uint32_t*>(arena + c.pflag_off + idx * kFlagSlotBytes);
}

// The `count` threads of role group `id` have issued their stores for chunk k: publish `seq` in every rank's slot.
// Thread t (< world) pairs with rank (rank + t) % world and writes word [my rank] of that rank's slot.
__device__ __forceinline__ void group_signal(const CommDev& c, int id, int count, int t, int kind, int k, uint32_t seq) {
  group_sync(id, count);  // every store of the group is ordered before the release below
  if (t < c.world) dev::st_release_sys(flag_slot(dev::peer_sel(c, t), c, kind, k) + c.rank, seq);
}
// Wait until every rank has published `seq` for (kind, k): thread t (< world) polls word [t] of my own slot.
__device__ __forceinline__ void group_wait(const CommDev& c, int id, int count, int t, int kind, int k, uint32_t seq) {
  if (t < c.world) dev::wait_flag(c, flag_slot(c.peer[0], c, kind, k) + t, seq);
  group_sync(id, count);  // peers' data is now visible to every thread of the group (the acquire invalidated L1)
}

}  // namespace pl

template <int MODE, int W, int ALG>
__global__ void __launch_bounds__(kThreads, 1)
    k_pipe(CommDev c, void* buf, unsigned long long n, float scale, int K, unsigned long long cell) {
  using namespace dev;
  using namespace pl;
  constexpr int WVB = Wire<MODE>::kBytes;
  constexpr int U = Unroll<W>::kU;
  const uint32_t seq0 = op_begin(c);
  const uint32_t seq = seq0 * 4u + 1u;
  const unsigned long long stage = (seq0 & 1u) ? c.stage_off[1] : c.stage_off[0];
  const bool aligned = buf_aligned<MODE>(buf);
  const unsigned long long V = (n + 7) / 8;
  const unsigned long long Ls = (V + W - 1) / W;
  const unsigned long long g = gridDim.x, b = blockIdx.x;
  const unsigned long long reduced = stage + static_cast<unsigned long long>(W) * c.slice_cap;  // P2P only
  uint8_t* const mine = c.peer[0];
  const int tid = threadIdx.x;

  if (tid < kA) {
    // ================= role A: cast (NVLS) / scatter (P2P) =================
    const int t = tid;
    if (t == 0) trace_stamp(c, 0);
    for (int k = 0; k < K; ++k) {
      const unsigned long long lo = (static_cast<unsigned long long>(k) * g + b) * cell;
      if (lo >= Ls) break;
      const unsigned long long hi = lo + cell < Ls ? lo + cell : Ls;
      for (unsigned long long v0 = lo + t; v0 < hi; v0 += static_cast<unsigned long long>(kA) * U) {
        F8 x[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kA;
#pragma unroll
          for (int jj = 0; jj < W; ++jj) {
            int j = c.rank + jj;
            if (j >= W) j -= W;
            const unsigned long long gv = j * Ls + v;
            if (v < hi && gv < V) x[u][jj] = load_in<MODE>(buf, gv * 8, n, aligned);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kA;
#pragma unroll
          for (int jj = 0; jj < W; ++jj) {
            int j = c.rank + jj;
            if (j >= W) j -= W;
            const unsigned long long gv = j * Ls + v;
            if (v < hi && gv < V) {
              const Wire<MODE> w = compress<MODE>(x[u][jj], scale);
              if constexpr (ALG == kNvls)
                st_wire<MODE>(mine + stage + gv * WVB, w);  // my stage: one message-sized buffer, vec gv at gv
              else
                st_wire<MODE>(c.peer[jj] + stage + c.rank * c.slice_cap + v * WVB, w);  // rank j's recv[me]
            }
          }
        }
      }
      group_signal(c, 1, kA, t, 0, k, seq);
    }
    if (t == 0) trace_stamp(c, 1);
  } else if (tid < kA + kB) {
    // ================= role B: move (NVLS) / reduce (P2P) =================
    const int t = tid - kA;
    const unsigned long long base = c.rank * Ls;  // my slice
    for (int k = 0; k < K; ++k) {
      const unsigned long long lo = (static_cast<unsigned long long>(k) * g + b) * cell;
      if (lo >= Ls) break;
      const unsigned long long hi = lo + cell < Ls ? lo + cell : Ls;
      group_wait(c, 2, kB, t, 0, k, seq);
      if (t == 0 && k == 0) trace_stamp(c, 2);
      if constexpr (ALG == kNvls) {
        constexpr int UM = MODE == B2_F32 ? 4 : 8;  // 128 B of switch-side reductions in flight per thread
        uint8_t* const mcs = c.mc + stage;
        for (unsigned long long v0 = lo + t; v0 < hi; v0 += static_cast<unsigned long long>(kB) * UM) {
          Wire<MODE> q[UM];
#pragma unroll
          for (int u = 0; u < UM; ++u) {
            const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kB;
            if (v < hi && base + v < V) q[u] = mm_ld_reduce_wire<MODE>(mcs + (base + v) * WVB);
          }
#pragma unroll
          for (int u = 0; u < UM; ++u) {
            const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kB;
            if (v < hi && base + v < V) mm_st_wire<MODE>(mcs + (base + v) * WVB, q[u]);
          }
        }
      } else {
        for (unsigned long long v0 = lo + t; v0 < hi; v0 += static_cast<unsigned long long>(kB) * U) {
          Wire<MODE> w[U][W];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kB;
            if (v < hi && base + v < V) {
#pragma unroll
              for (int r = 0; r < W; ++r) w[u][r] = ld_wire<MODE>(mine + stage + r * c.slice_cap + v * WVB);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kB;
            if (v < hi && base + v < V) {
              F8 s = widen<MODE>(w[u][0]);
#pragma unroll
              for (int r = 1; r < W; ++r) accumulate(s, widen<MODE>(w[u][r]));  // rank order, fp32
              st_wire<MODE>(mine + reduced + v * WVB, finalize<MODE>(s));
            }
          }
        }
      }
      group_signal(c, 2, kB, t, 1, k, seq);
    }
    if (t == 0) trace_stamp(c, 3);
  } else {
    // ================= role C: widen (NVLS) / gather (P2P) =================
    const int t = tid - kA - kB;
    for (int k = 0; k < K; ++k) {
      const unsigned long long lo = (static_cast<unsigned long long>(k) * g + b) * cell;
      if (lo >= Ls) break;
      const unsigned long long hi = lo + cell < Ls ? lo + cell : Ls;
      group_wait(c, 3, kC, t, 1, k, seq);
      if (t == 0 && k == 0) trace_stamp(c, 4);
      for (unsigned long long v0 = lo + t; v0 < hi; v0 += static_cast<unsigned long long>(kC) * U) {
        Wire<MODE> w[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kC;
#pragma unroll
          for (int jj = 0; jj < W; ++jj) {
            int j = c.rank + jj;
            if (j >= W) j -= W;
            const unsigned long long gv = j * Ls + v;
            if (v < hi && gv < V) {
              if constexpr (ALG == kNvls)
                w[u][jj] = ld_wire<MODE>(mine + stage + gv * WVB);  // the switch replicated every slice into my stage
              else
                w[u][jj] = ld_wire<MODE>(c.peer[jj] + reduced + v * WVB);  // rank j's reduced slice, over NVLink
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned long long v = v0 + static_cast<unsigned long long>(u) * kC;
#pragma unroll
          for (int jj = 0; jj < W; ++jj) {
            int j = c.rank + jj;
            if (j >= W) j -= W;
            const unsigned long long gv = j * Ls + v;
            if (v < hi && gv < V) store_out<MODE>(buf, gv * 8, n, aligned, w[u][jj]);
          }
        }
      }
    }
    if (t == 0) trace_stamp(c, 5);
  }
  op_end(c, seq0);
}
