// b200ddp.cu — libb200ddp.so: B200 (sm_100a) data plane for the `local_cuda` TorchX scheduler.
//
// What lives here (see include/b200ddp.h for the ABI and DESIGN.md for the rationale):
//   * rendezvous: POSIX-shm control block + CUDA-IPC exchange of ONE symmetric arena per rank
//     (replaces c10d TCPStore + ncclCommInitRank on the reference path,
//      torchx/distributed/__init__.py:217-222 -> torch.distributed.init_process_group)
//   * the DDP gradient-bucket allreduce as ONE fused kernel per bucket
//     (replaces the 4-launch cast -> div -> ncclAllReduce -> copy sequence of
//      torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:57-93)
//       - one-shot  : push compressed message to every peer, one flag barrier, reduce locally
//       - two-shot  : push-scatter (fp32->wire cast + scale fused into the NVLink stores)
//                     -> reduce own slice (fp32 accumulate, rank order) -> pull-gather
//                     (wire->fp32 cast fused into the NVLink loads)
//   * broadcast / barrier on the same fabric (DDP init + BN-buffer sync, dist.barrier()).
//
// Memory model: every cross-GPU hand-off is  data stores -> bar.sync -> st.release.sys(flag)
// on the producer and  ld.acquire.sys(flag) -> bar.sync -> data loads  on the consumer, with a
// monotonically increasing sequence number instead of flag resets (no ABA, no reset races).
// All peer waits are bounded: a CTA that waits longer than `timeout_ns` records B2_ETIMEOUT in a
// host-mapped status word and carries on, so a dead peer can never hang the GPU.
//
// No tensor cores: the path is a pure bandwidth-bound reduction (1 add per 2-4 bytes moved).

#include "../../include/b200ddp.h"

#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <new>
#include <string>

// ------------------------------------------------------------------------------------------------
// constants shared by host and device
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kThreads = 512;                // threads per CTA for every kernel in this file
constexpr int kMaxCtas = 296;                // 2 x 148 SMs: upper bound on the grid of a collective
constexpr int kFlagSlotBytes = 32;           // one 32 B sector of flags per CTA index (8 x u32, one per peer)
constexpr size_t kFlagRegionBytes = 64 << 10;  // >= kMaxCtas * kFlagSlotBytes, keeps stages 64 KiB aligned
constexpr size_t kDefaultStageBytes = 512ull << 20;  // x2 stages = 1 GiB of the 180 GB: a 1 GiB fp32 bucket is one launch
constexpr unsigned long long kDefaultTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;

static_assert(kMaxCtas * kFlagSlotBytes <= (int)kFlagRegionBytes, "flag region too small");

// Device-visible description of one rank's view of the communicator; passed BY VALUE as a kernel
// parameter (well under the 4 KiB parameter limit), so no device-side indirection is needed.
struct CommDev {
  int rank;
  int world;
  uint8_t* peer[B2_MAX_WORLD];      // peer[jj] = symmetric arena of rank (rank + jj) % world as mapped in
                                    // THIS process (peer[0] is this rank's own).  Pre-rotated on the host so
                                    // unrolled device loops index it with compile-time constants (registers,
                                    // not a local-memory copy of the parameter block) and so the W ranks
                                    // never all target the same peer in the same loop step.
  uint8_t* abs[B2_MAX_WORLD];       // abs[r] = rank r's arena (absolute order): used where values must be combined
                                    // in rank order, so the loop index is both the pointer index and the rank
  uint32_t* opseq;                  // local: number of collectives completed on this communicator
  uint32_t* done;                   // local: CTAs of the running collective that reached the epilogue
  uint32_t* status;                 // host-mapped: 0 = healthy, else a B2_E* code (positive)
  unsigned long long timeout_ns;    // bound on any single peer wait
  unsigned long long flag_off;      // byte offset of this LANE's flag region inside an arena
  unsigned long long stage_off[2];  // byte offsets of the two staging buffers inside an arena
  unsigned long long slice_cap;     // bytes of one region; a stage is (world + 1) regions:
                                    //   regions 0..W-1 = "recv[r]" (written by rank r), region W = "reduced"
  uint32_t* stagger_ctr;            // local: number of split collectives whose lane-0 half has finished its scatter
  int stagger_role;                 // 0 = none, 1 = signal after my scatter (lane 0), 2 = wait before my scatter (lane 1)
  unsigned long long* trace;        // optional (b2_comm_trace): per-CTA globaltimer stamps of the LAST collective,
                                    // 8 slots per CTA: start, A done, bar1 done, B done, bar2 done, C done
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
namespace dev {

struct F8 {
  float v[8];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// 256-bit / 128-bit streaming accesses (LDG.E.NA.256 / STG.E.NA.256 on sm_100a).
__device__ __forceinline__ F8 ldg_f8(const float* p) {
  F8 r;
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]),
                 "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void stg_f8(float* p, const F8& r) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
               "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]),
               "f"(r.v[6]), "f"(r.v[7])
               : "memory");
}
__device__ __forceinline__ uint4 ldg_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void stg_u4(void* p, const uint4& r) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y),
               "r"(r.z), "r"(r.w)
               : "memory");
}

// fp32 pair -> packed bf16x2 with round-to-nearest-even (one F2FP.BF16.F32.PACK_AB). `lo` lands
// in bits [15:0] (the lower address in little-endian memory), `hi` in bits [31:16].
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// ---- per-mode traits ---------------------------------------------------------------------------
// A "vec" is 8 consecutive elements everywhere in this file.
template <int MODE>
struct Wire;  // wire representation of one vec

template <>
struct Wire<B2_F32_WIRE_BF16> {
  static constexpr int kBytes = 16;
  uint4 q;
};
template <>
struct Wire<B2_BF16> {
  static constexpr int kBytes = 16;
  uint4 q;
};
template <>
struct Wire<B2_F32> {
  static constexpr int kBytes = 32;
  F8 f;
};

template <int MODE>
__device__ __forceinline__ Wire<MODE> ld_wire(const uint8_t* p) {
  Wire<MODE> w;
  if constexpr (MODE == B2_F32) {
    w.f = ldg_f8(reinterpret_cast<const float*>(p));
  } else {
    w.q = ldg_u4(p);
  }
  return w;
}
template <int MODE>
__device__ __forceinline__ void st_wire(uint8_t* p, const Wire<MODE>& w) {
  if constexpr (MODE == B2_F32) {
    stg_f8(reinterpret_cast<float*>(p), w.f);
  } else {
    stg_u4(p, w.q);
  }
}

// wire(scale * x): the value a rank contributes.  Rounding points are part of the contract
// (oracle/allreduce_oracle.c: b2o_compress):
//   F32_WIRE_BF16 : bf16( float(bf16(x)) * scale )      == `buf.to(bf16).div_(W)` for W = 2^k
//   BF16          : bf16( float(x) * scale )            (x is already bf16)
//   F32           : x * scale                           == Reducer's `mul_out(bucket, grad, 1/W)`
template <int MODE>
__device__ __forceinline__ Wire<MODE> compress(const F8& x, float scale) {
  Wire<MODE> w;
  if constexpr (MODE == B2_F32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) w.f.v[i] = __fmul_rn(x.v[i], scale);
  } else {
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = x.v[2 * i], b = x.v[2 * i + 1];
      if constexpr (MODE == B2_F32_WIRE_BF16) {
        const uint32_t p = pack_bf16x2(a, b);  // first rounding: the `.to(bf16)` cast
        a = bf16_lo(p);
        b = bf16_hi(p);
      }
      o[i] = pack_bf16x2(__fmul_rn(a, scale), __fmul_rn(b, scale));
    }
    w.q = make_uint4(o[0], o[1], o[2], o[3]);
  }
  return w;
}

template <int MODE>
__device__ __forceinline__ F8 widen(const Wire<MODE>& w) {
  if constexpr (MODE == B2_F32) {
    return w.f;
  } else {
    F8 r;
    r.v[0] = bf16_lo(w.q.x);
    r.v[1] = bf16_hi(w.q.x);
    r.v[2] = bf16_lo(w.q.y);
    r.v[3] = bf16_hi(w.q.y);
    r.v[4] = bf16_lo(w.q.z);
    r.v[5] = bf16_hi(w.q.z);
    r.v[6] = bf16_lo(w.q.w);
    r.v[7] = bf16_hi(w.q.w);
    return r;
  }
}

// round(s): the reduced value as it travels in the gather phase / is stored.
template <int MODE>
__device__ __forceinline__ Wire<MODE> finalize(const F8& s) {
  Wire<MODE> w;
  if constexpr (MODE == B2_F32) {
    w.f = s;
  } else {
    w.q = make_uint4(pack_bf16x2(s.v[0], s.v[1]), pack_bf16x2(s.v[2], s.v[3]),
                     pack_bf16x2(s.v[4], s.v[5]), pack_bf16x2(s.v[6], s.v[7]));
  }
  return w;
}

__device__ __forceinline__ void accumulate(F8& s, const F8& c) {
#pragma unroll
  for (int i = 0; i < 8; ++i) s.v[i] = __fadd_rn(s.v[i], c.v[i]);
}

// ---- local bucket accesses (the caller's tensor: any alignment, any length) --------------------
template <int MODE>
__device__ __forceinline__ F8 load_in(const void* buf, unsigned long long e, unsigned long long n,
                                      bool aligned) {
  F8 x;
  if constexpr (MODE == B2_BF16) {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(buf) + e;
    if (aligned && e + 8 <= n) {
      Wire<B2_BF16> w;
      w.q = ldg_u4(p);
      x = widen<B2_BF16>(w);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x.v[i] = (e + i < n) ? __uint_as_float(static_cast<uint32_t>(p[i]) << 16) : 0.f;
    }
  } else {
    const float* p = reinterpret_cast<const float*>(buf) + e;
    if (aligned && e + 8 <= n) {
      x = ldg_f8(p);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x.v[i] = (e + i < n) ? p[i] : 0.f;
    }
  }
  return x;
}

// `w` is the reduced vec in wire format; writes it to the caller's tensor in the tensor's dtype.
template <int MODE>
__device__ __forceinline__ void store_out(void* buf, unsigned long long e, unsigned long long n,
                                          bool aligned, const Wire<MODE>& w) {
  if constexpr (MODE == B2_BF16) {
    uint16_t* p = reinterpret_cast<uint16_t*>(buf) + e;
    if (aligned && e + 8 <= n) {
      stg_u4(p, w.q);
    } else {
      const uint32_t q[4] = {w.q.x, w.q.y, w.q.z, w.q.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (e + i < n) p[i] = static_cast<uint16_t>((q[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
    }
  } else {
    float* p = reinterpret_cast<float*>(buf) + e;
    const F8 r = widen<MODE>(w);
    if (aligned && e + 8 <= n) {
      stg_f8(p, r);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (e + i < n) p[i] = r.v[i];
    }
  }
}

template <int MODE>
__device__ __forceinline__ bool buf_aligned(const void* buf) {
  return (reinterpret_cast<uintptr_t>(buf) & (MODE == B2_BF16 ? 15u : 31u)) == 0;
}

// ---- cross-GPU barrier among the CTAs with the same blockIdx.x on every rank --------------------
// Thread p (< world) publishes `seq` into peer p's flag slot for this CTA index and waits for peer
// p's `seq` in its own slot.  Sequence numbers only grow, so "flag >= seq" (wrap-safe) is the test.
__device__ __forceinline__ void cta_xbar(const CommDev& c, uint32_t seq) {
  __syncthreads();  // all of this CTA's data stores are ordered before the release below
  if (threadIdx.x < c.world) {
    const int jj = threadIdx.x;  // this thread pairs with rank p = (rank + jj) % world
    int p = c.rank + jj;
    if (p >= c.world) p -= c.world;
    uint8_t* their_arena = c.peer[0];
#pragma unroll
    for (int i = 1; i < B2_MAX_WORLD; ++i)
      if (jj == i) their_arena = c.peer[i];  // select chain: keeps the parameter block out of local memory
    const size_t slot = c.flag_off + static_cast<size_t>(blockIdx.x) * kFlagSlotBytes;
    uint32_t* theirs = reinterpret_cast<uint32_t*>(their_arena + slot) + c.rank;
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(c.peer[0] + slot) + p;
    st_release_sys(theirs, seq);
    unsigned long long t0 = 0;
    unsigned spins = 0;
    // Poll with ld.acquire.sys itself.  The alternative (relaxed polling + one fence.acq_rel.sys at the end) was measured
    // on 4xB200 and DOUBLED the cost of the second barrier (5 -> 11 us): the standalone fence is a full MEMBAR.SYS that
    // also drains this SM's outstanding stores, the acquire load is not.
    while (static_cast<int32_t>(ld_acquire_sys(mine) - seq) < 0) {
      if ((++spins & 63u) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) {
          t0 = now;
        } else if (now - t0 > c.timeout_ns) {
          *reinterpret_cast<volatile uint32_t*>(c.status) = static_cast<uint32_t>(-B2_ETIMEOUT);
          __threadfence_system();
          break;  // give up: results are undefined, but the GPU is not hung
        }
      }
    }
  }
  __syncthreads();  // peers' data is now visible to every thread of this CTA
}

// Every collective kernel starts by reading the communicator's op counter (parity selects the
// staging buffer, the value seeds this op's flag sequence numbers) and ends by bumping it once
// all CTAs are through.  Keeping the counter on the device makes the launch sequence CUDA-graph
// replayable and keeps the host stateless.
__device__ __forceinline__ uint32_t op_begin(const CommDev& c) { return ld_volatile_u32(c.opseq); }

__device__ __forceinline__ void op_end(const CommDev& c, uint32_t seq0) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(c.done, 1u) == gridDim.x - 1) {
      *reinterpret_cast<volatile uint32_t*>(c.done) = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(c.opseq) = seq0 + 1;
    }
  }
}

// Staggered lanes: the lane-1 half of a split collective starts its scatter only when the lane-0 half has issued its own,
// so each half's barriers fall into the other half's NVLink phases instead of coinciding with them.
__device__ __forceinline__ void stagger_signal(const CommDev& c) {
  if (c.stagger_role == 1 && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(c.stagger_ctr, 1u);
}
__device__ __forceinline__ void stagger_wait(const CommDev& c, uint32_t seq0) {
  if (c.stagger_role != 2) return;
  if (threadIdx.x == 0) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (static_cast<int32_t>(ld_volatile_u32(c.stagger_ctr) - (seq0 + 1u)) < 0) {
      if ((++spins & 63u) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > c.timeout_ns) break;  // lane 0 never came: proceed, the peer barriers will report it
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void trace_stamp(const CommDev& c, int slot) {
  if (c.trace != nullptr && threadIdx.x == 0) c.trace[blockIdx.x * 8 + slot] = globaltimer_ns();
}

template <int W>
struct Unroll {  // vecs per thread per loop trip, chosen so ~8 wire vecs are in flight per thread
  static constexpr int kU = (W >= 8) ? 1 : (W >= 4 ? 2 : (W >= 2 ? 4 : 8));
};

}  // namespace dev

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// W == 1 (and the single-GPU roofline probe): x <- round(wire(scale * x)), one streaming pass.
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_local_pass(void* buf, unsigned long long n, float scale) {
  using namespace dev;
  const bool aligned = buf_aligned<MODE>(buf);
  const unsigned long long V = (n + 7) / 8;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kThreads;
  constexpr int U = 4;
  for (unsigned long long v0 = static_cast<unsigned long long>(blockIdx.x) * kThreads + threadIdx.x;
       v0 < V; v0 += stride * U) {
    F8 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
      if (v < V) x[u] = load_in<MODE>(buf, v * 8, n, aligned);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
      if (v < V) {
        const Wire<MODE> c = compress<MODE>(x[u], scale);
        store_out<MODE>(buf, v * 8, n, aligned, finalize<MODE>(widen<MODE>(c)));
      }
    }
  }
}

// ---- TMA-staged variant of the local pass --------------------------------------------------------------------
// Persistent CTAs stream 16 KiB tiles through a 4-deep shared-memory ring: one elected thread issues
// cp.async.bulk (global -> shared, completion on an mbarrier; SASS UBLKCP), all threads round the tile in place in
// shared memory, and the tile goes back with a bulk store (shared -> global, bulk_group).  Loads of the next tiles are
// always in flight while the current tile is being rounded and stored, with no registers tied up by outstanding
// loads - 64 KiB of HBM reads in flight per CTA from a single issuing thread.
namespace tma {

constexpr int kTileBytes = 16 << 10;
constexpr int kStages = 4;
constexpr int kTmaThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// x <- round(wire(scale * x)) on one 16-byte group (4 fp32 or 8 bf16), same arithmetic as compress+finalize.
template <int MODE>
__device__ __forceinline__ uint4 round16(uint4 q, float scale) {
  using namespace dev;
  if constexpr (MODE == B2_BF16) {
    uint32_t in[4] = {q.x, q.y, q.z, q.w}, out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = pack_bf16x2(__fmul_rn(bf16_lo(in[i]), scale), __fmul_rn(bf16_hi(in[i]), scale));
    return make_uint4(out[0], out[1], out[2], out[3]);
  } else {
    float f[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      if constexpr (MODE == B2_F32) {
        o[i] = __float_as_uint(__fmul_rn(f[i], scale));
        o[i + 1] = __float_as_uint(__fmul_rn(f[i + 1], scale));
      } else {
        const uint32_t p = pack_bf16x2(f[i], f[i + 1]);                                        // .to(bf16)
        const uint32_t r = pack_bf16x2(__fmul_rn(bf16_lo(p), scale), __fmul_rn(bf16_hi(p), scale));  // .div_(W), bf16
        o[i] = r << 16;             // widen back to fp32: bf16 bits in the high half
        o[i + 1] = r & 0xffff0000u;
      }
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace tma

template <int MODE>
__global__ void __launch_bounds__(tma::kTmaThreads) k_local_pass_tma(void* buf, unsigned long long n, float scale) {
  using namespace tma;
  extern __shared__ __align__(128) uint8_t ring_raw[];  // kStages * kTileBytes of dynamic shared memory
  uint8_t(*ring)[kTileBytes] = reinterpret_cast<uint8_t(*)[kTileBytes]>(ring_raw);
  __shared__ alignas(8) uint64_t full[kStages];
  constexpr int kElem = MODE == B2_BF16 ? 2 : 4;
  const unsigned long long bytes = n * kElem;
  const unsigned long long ntiles = bytes / kTileBytes;  // full tiles go through TMA; the tail is handled below
  uint8_t* base = static_cast<uint8_t*>(buf);
  const unsigned long long my_first = blockIdx.x;
  const unsigned long long step = gridDim.x;
  const unsigned long long my_count = my_first < ntiles ? (ntiles - my_first + step - 1) / step : 0;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      if (static_cast<unsigned long long>(s) < my_count) {
        mbar_arrive_expect_tx(&full[s], kTileBytes);
        bulk_g2s(ring[s], base + (my_first + s * step) * kTileBytes, kTileBytes, &full[s]);
      }
    }
  }
  for (unsigned long long k = 0; k < my_count; ++k) {
    const int s = static_cast<int>(k % kStages);
    mbar_wait(&full[s], static_cast<uint32_t>((k / kStages) & 1));
    uint4* tile = reinterpret_cast<uint4*>(ring[s]);
#pragma unroll
    for (int i = 0; i < kTileBytes / 16 / kTmaThreads; ++i) {
      const int idx = i * kTmaThreads + threadIdx.x;  // conflict-free: consecutive lanes, consecutive 16 B
      tile[idx] = round16<MODE>(tile[idx], scale);
    }
    fence_proxy_async();  // my generic-proxy writes to the tile are visible to the bulk store (async proxy)
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_s2g(base + (my_first + k * step) * kTileBytes, ring[s], kTileBytes);
      bulk_commit();
      // Refill the stage whose store was issued ONE iteration ago: allowing one group in flight means that older
      // store has finished reading shared memory, while the store just issued keeps draining.
      if (k >= 1 && k - 1 + kStages < my_count) {
        bulk_wait_read<1>();
        const int sp = static_cast<int>((k - 1) % kStages);
        mbar_arrive_expect_tx(&full[sp], kTileBytes);
        bulk_g2s(ring[sp], base + (my_first + (k - 1 + kStages) * step) * kTileBytes, kTileBytes, &full[sp]);
      }
    }
  }
  if (threadIdx.x == 0) bulk_wait_read<0>();  // shared memory must outlive the last store's reads
  // tail (< 16 KiB): plain loads/stores, spread over the grid
  const unsigned long long tail0 = ntiles * kTileBytes / kElem;
  for (unsigned long long e = tail0 + (static_cast<unsigned long long>(blockIdx.x) * kTmaThreads + threadIdx.x) * 8; e < n;
       e += static_cast<unsigned long long>(gridDim.x) * kTmaThreads * 8) {
    const dev::F8 x = dev::load_in<MODE>(buf, e, n, false);
    const dev::Wire<MODE> c = dev::compress<MODE>(x, scale);
    dev::store_out<MODE>(buf, e, n, false, dev::finalize<MODE>(dev::widen<MODE>(c)));
  }
}

// One-shot: latency regime.  Wire traffic per rank: (W-1) * S out, (W-1) * S in.
template <int MODE, int W>
__global__ void __launch_bounds__(kThreads, 1)
    k_oneshot(CommDev c, void* buf, unsigned long long n, float scale) {
  using namespace dev;
  constexpr int WVB = Wire<MODE>::kBytes;
  constexpr int U = Unroll<W>::kU;
  const uint32_t seq0 = op_begin(c);
  const unsigned long long stage = (seq0 & 1u) ? c.stage_off[1] : c.stage_off[0];
  const bool aligned = buf_aligned<MODE>(buf);
  const unsigned long long V = (n + 7) / 8;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kThreads;
  const unsigned long long first = static_cast<unsigned long long>(blockIdx.x) * kThreads + threadIdx.x;
  trace_stamp(c, 0);

  // phase A: compress my message once, push it into recv[rank] of every rank (mine included)
  for (unsigned long long v0 = first; v0 < V; v0 += stride * U) {
    F8 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
      if (v < V) x[u] = load_in<MODE>(buf, v * 8, n, aligned);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
      if (v < V) {
        const Wire<MODE> w = compress<MODE>(x[u], scale);
#pragma unroll
        for (int jj = 0; jj < W; ++jj)  // peer[] is rotated: the W ranks never hammer one peer at a time
          st_wire<MODE>(c.peer[jj] + stage + c.rank * c.slice_cap + v * WVB, w);
      }
    }
  }
  trace_stamp(c, 1);
  cta_xbar(c, seq0 * 4u + 1u);
  trace_stamp(c, 2);

  // phase B: reduce the W messages (all local now) in rank order, write the caller's tensor
  const uint8_t* mine = c.peer[0] + stage;
  for (unsigned long long v0 = first; v0 < V; v0 += stride * U) {
    Wire<MODE> w[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
      if (v < V) {
#pragma unroll
        for (int r = 0; r < W; ++r) w[u][r] = ld_wire<MODE>(mine + r * c.slice_cap + v * WVB);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
      if (v < V) {
        F8 s = widen<MODE>(w[u][0]);
#pragma unroll
        for (int r = 1; r < W; ++r) accumulate(s, widen<MODE>(w[u][r]));
        store_out<MODE>(buf, v * 8, n, aligned, finalize<MODE>(s));
      }
    }
  }
  trace_stamp(c, 3);
  op_end(c, seq0);
}

// Two-shot: bandwidth regime.  The message is cut into W slices of Ls vecs; rank i owns slice i.
//   A  push-scatter : read my fp32 bucket once, cast+scale, STORE slice j into rank j's recv[me]
//   B  reduce       : sum recv[0..W-1] of my slice (local HBM), fp32 accumulate in rank order,
//                     round once, write my "reduced" region
//   C  pull-gather  : LOAD slice j from rank j's "reduced" region over NVLink, widen, write bucket
// Wire traffic per rank and direction: 2 * (W-1)/W * S  (the allreduce lower bound for P2P).
// CTA b touches the same vec indices of a slice on every rank and in every phase, so the only
// synchronisation needed is among the CTAs with equal blockIdx.x across ranks (no grid sync).
template <int MODE, int W>
__global__ void __launch_bounds__(kThreads, 1)
    k_twoshot(CommDev c, void* buf, unsigned long long n, float scale) {
  using namespace dev;
  constexpr int WVB = Wire<MODE>::kBytes;
  constexpr int U = Unroll<W>::kU;
  const uint32_t seq0 = op_begin(c);
  const unsigned long long stage = (seq0 & 1u) ? c.stage_off[1] : c.stage_off[0];
  const bool aligned = buf_aligned<MODE>(buf);
  const unsigned long long V = (n + 7) / 8;
  const unsigned long long Ls = (V + W - 1) / W;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kThreads;
  const unsigned long long first = static_cast<unsigned long long>(blockIdx.x) * kThreads + threadIdx.x;
  const unsigned long long my_recv = stage + c.rank * c.slice_cap;
  const unsigned long long reduced = stage + static_cast<unsigned long long>(W) * c.slice_cap;
  stagger_wait(c, seq0);
  trace_stamp(c, 0);

  // ---- phase A -------------------------------------------------------------------------------
  for (unsigned long long v0 = first; v0 < Ls; v0 += stride * U) {
    F8 x[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = (c.rank + jj) % W;
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V) x[u][jj] = load_in<MODE>(buf, gv * 8, n, aligned);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = (c.rank + jj) % W;
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V)
          st_wire<MODE>(c.peer[jj] + my_recv + v * WVB, compress<MODE>(x[u][jj], scale));
      }
    }
  }
  trace_stamp(c, 1);
  stagger_signal(c);
  cta_xbar(c, seq0 * 4u + 1u);
  trace_stamp(c, 2);

  // ---- phase B -------------------------------------------------------------------------------
  {
    uint8_t* mine = c.peer[0];
    const unsigned long long base = c.rank * Ls;
    for (unsigned long long v0 = first; v0 < Ls; v0 += stride * U) {
      Wire<MODE> w[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned long long v = v0 + u * stride;
        if (v < Ls && base + v < V) {
#pragma unroll
          for (int r = 0; r < W; ++r)
            w[u][r] = ld_wire<MODE>(mine + stage + r * c.slice_cap + v * WVB);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned long long v = v0 + u * stride;
        if (v < Ls && base + v < V) {
          F8 s = widen<MODE>(w[u][0]);
#pragma unroll
          for (int r = 1; r < W; ++r) accumulate(s, widen<MODE>(w[u][r]));
          st_wire<MODE>(mine + reduced + v * WVB, finalize<MODE>(s));
        }
      }
    }
  }
  trace_stamp(c, 3);
  cta_xbar(c, seq0 * 4u + 2u);
  trace_stamp(c, 4);

  // ---- phase C -------------------------------------------------------------------------------
  for (unsigned long long v0 = first; v0 < Ls; v0 += stride * U) {
    Wire<MODE> w[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = (c.rank + jj) % W;
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V) w[u][jj] = ld_wire<MODE>(c.peer[jj] + reduced + v * WVB);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = (c.rank + jj) % W;
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V) store_out<MODE>(buf, gv * 8, n, aligned, w[u][jj]);
      }
    }
  }
  trace_stamp(c, 5);
  op_end(c, seq0);
}

// Two-shot, pull-pull variant.  Measured on 8xB200 (profiles/r01_phase_trace_w4.md): remote LOADS stream at the full
// NVLink rate (~780 GB/s per GPU) while the push-scatter of k_twoshot only reaches ~440 GB/s once the release fence has
// drained its stores, and a barrier that follows purely LOCAL stores costs ~3 us instead of 7-17 us.  So here nothing is
// ever stored to a peer except flags:
//   A0 compress : read my bucket once, cast+scale, store ALL W slices into MY OWN stage (L2-resident for DDP bucket sizes)
//   B  pull-reduce : LOAD slice `me` from every rank's stage (rank order), fp32 accumulate, round once -> my "reduced"
//   C  pull-gather : as in k_twoshot
// The separate local reduce pass disappears (it is fused into the loads of B).
template <int MODE, int W>
__global__ void __launch_bounds__(kThreads, 1)
    k_twoshot_pull(CommDev c, void* buf, unsigned long long n, float scale) {
  using namespace dev;
  constexpr int WVB = Wire<MODE>::kBytes;
  constexpr int U = Unroll<W>::kU;
  const uint32_t seq0 = op_begin(c);
  const unsigned long long stage = (seq0 & 1u) ? c.stage_off[1] : c.stage_off[0];
  const bool aligned = buf_aligned<MODE>(buf);
  const unsigned long long V = (n + 7) / 8;
  const unsigned long long Ls = (V + W - 1) / W;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kThreads;
  const unsigned long long first = static_cast<unsigned long long>(blockIdx.x) * kThreads + threadIdx.x;
  const unsigned long long reduced = stage + static_cast<unsigned long long>(W) * c.slice_cap;
  uint8_t* mine = c.peer[0];
  trace_stamp(c, 0);

  // ---- phase A0: region j of my stage <- my compressed slice j ------------------------------------
  for (unsigned long long v0 = first; v0 < Ls; v0 += stride * U) {
    F8 x[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V) x[u][j] = load_in<MODE>(buf, gv * 8, n, aligned);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V)
          st_wire<MODE>(mine + stage + j * c.slice_cap + v * WVB, compress<MODE>(x[u][j], scale));
      }
    }
  }
  trace_stamp(c, 1);
  cta_xbar(c, seq0 * 4u + 1u);
  trace_stamp(c, 2);

  // ---- phase B: pull my slice from every rank, reduce in rank order --------------------------------
  {
    const unsigned long long base = c.rank * Ls;
    const unsigned long long my_region = stage + c.rank * c.slice_cap;
    for (unsigned long long v0 = first; v0 < Ls; v0 += stride * U) {
      Wire<MODE> w[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned long long v = v0 + u * stride;
        if (v < Ls && base + v < V) {
#pragma unroll
          for (int r = 0; r < W; ++r) w[u][r] = ld_wire<MODE>(c.abs[r] + my_region + v * WVB);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned long long v = v0 + u * stride;
        if (v < Ls && base + v < V) {
          F8 s = widen<MODE>(w[u][0]);
#pragma unroll
          for (int r = 1; r < W; ++r) accumulate(s, widen<MODE>(w[u][r]));
          st_wire<MODE>(mine + reduced + v * WVB, finalize<MODE>(s));
        }
      }
    }
  }
  trace_stamp(c, 3);
  cta_xbar(c, seq0 * 4u + 2u);
  trace_stamp(c, 4);

  // ---- phase C: pull-gather ---------------------------------------------------------------------
  for (unsigned long long v0 = first; v0 < Ls; v0 += stride * U) {
    Wire<MODE> w[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = (c.rank + jj) % W;
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V) w[u][jj] = ld_wire<MODE>(c.peer[jj] + reduced + v * WVB);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long v = v0 + u * stride;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int j = (c.rank + jj) % W;
        const unsigned long long gv = j * Ls + v;
        if (v < Ls && gv < V) store_out<MODE>(buf, gv * 8, n, aligned, w[u][jj]);
      }
    }
  }
  trace_stamp(c, 5);
  op_end(c, seq0);
}

// Broadcast of raw bytes: root pushes into every peer's stage, one barrier, peers copy out.
__global__ void __launch_bounds__(kThreads, 1)
    k_broadcast(CommDev c, uint8_t* buf, unsigned long long bytes, int root) {
  using namespace dev;
  const uint32_t seq0 = op_begin(c);
  const unsigned long long stage = (seq0 & 1u) ? c.stage_off[1] : c.stage_off[0];
  const bool aligned = (reinterpret_cast<uintptr_t>(buf) & 15u) == 0;
  const unsigned long long nvec = aligned ? bytes / 16 : 0;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * kThreads;
  const unsigned long long first = static_cast<unsigned long long>(blockIdx.x) * kThreads + threadIdx.x;
  if (c.rank == root) {
    for (unsigned long long v = first; v < nvec; v += stride) {
      const uint4 q = ldg_u4(buf + v * 16);
#pragma unroll
      for (int jj = 1; jj < B2_MAX_WORLD; ++jj)
        if (jj < c.world) stg_u4(c.peer[jj] + stage + v * 16, q);
    }
    for (unsigned long long b = nvec * 16 + first; b < bytes; b += stride) {
      const uint8_t x = buf[b];
#pragma unroll
      for (int jj = 1; jj < B2_MAX_WORLD; ++jj)
        if (jj < c.world) c.peer[jj][stage + b] = x;
    }
  }
  cta_xbar(c, seq0 * 4u + 1u);
  if (c.rank != root) {
    const uint8_t* src = c.peer[0] + stage;
    for (unsigned long long v = first; v < nvec; v += stride) stg_u4(buf + v * 16, ldg_u4(src + v * 16));
    for (unsigned long long b = nvec * 16 + first; b < bytes; b += stride) buf[b] = src[b];
  }
  op_end(c, seq0);
}

__global__ void __launch_bounds__(kThreads, 1) k_barrier(CommDev c) {
  using namespace dev;
  const uint32_t seq0 = op_begin(c);
  cta_xbar(c, seq0 * 4u + 1u);
  op_end(c, seq0);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char tmp[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tmp, sizeof(tmp), fmt, ap);
  va_end(ap);
  g_err = tmp;
  return code;
}

#define B2_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(B2_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                  \
  } while (0)

struct DeviceGuard {  // the library never leaves the caller's current device changed
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// ---- shm control block (multi-process rendezvous) ----------------------------------------------
constexpr uint64_t kShmMagic = 0x42323030444450ull;  // "B200DDP"

struct ShmSlot {
  cudaIpcMemHandle_t handle;
  int device;        // the rank's ordinal in ITS OWN numbering (CUDA_VISIBLE_DEVICES may differ between ranks)
  char bus_id[24];   // PCI bus id: the identity that is comparable across processes
  int pid;
  unsigned long long arena_bytes;
  std::atomic<uint32_t> ready;  // 1 once handle/device/pid are valid
  char pad[64];
};

struct ShmBlock {
  std::atomic<uint64_t> magic;
  uint64_t epoch;
  int world;
  std::atomic<int> mapped;    // ranks that have opened every peer handle
  std::atomic<int> departed;  // ranks that have finished using peer memory (destroy handshake)
  ShmSlot slot[B2_MAX_WORLD];
};

}  // namespace

// A communicator has two LANES: fully independent flag regions, stages and op counters inside the same arena.  Every
// collective runs on lane 0; a large two-shot allreduce is split in two halves that run CONCURRENTLY, lane 1 on an
// internal stream forked from / joined into the caller's stream, so that the barriers, the local reduce and the launch
// gap of one half overlap the NVLink phases of the other (profiles/r01_phase_trace_w4.md: ~25 us of a 55 us bucket
// collective is not wire time).
constexpr int kLanes = 2;

struct b2_comm {
  CommDev d{};        // lane 0
  CommDev d1{};       // lane 1 (same peers; its own flags / stages / counters)
  cudaStream_t lane_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  size_t split_min_wire_bytes = 0;  // two-shot messages of at least this many wire bytes are split over the lanes
  bool stagger = true;              // B2_STAGGER=0: start both halves together (the symmetric split)
  int device = -1;
  bool local_world = false;    // created by b2_comm_create_local (no IPC, no shm)
  bool peer_is_ipc[B2_MAX_WORLD] = {};
  uint8_t* arena_of[B2_MAX_WORLD] = {};  // arena_of[r] = rank r's arena as mapped in this process
  void* arena = nullptr;       // cudaMalloc'ed: [flags | stage0 | stage1]
  size_t arena_bytes = 0;
  size_t stage_bytes = 0;
  uint32_t* counters = nullptr;  // cudaMalloc'ed: opseq, done
  unsigned long long* trace_dev = nullptr;  // cudaMalloc'ed on demand: kMaxCtas * 8 stamps
  uint32_t* status_host = nullptr;
  int max_ctas = 0;              // 0 = heuristic
  int auto_twoshot = B2_ALGO_TWOSHOT;  // which two-shot AUTO uses (B2_AUTO_TWOSHOT=2|3 overrides)
  size_t oneshot_max_wire_bytes = 0;  // 0 = per-world default (see default_oneshot_max); B2_ONESHOT_MAX_BYTES overrides
  uint64_t launches = 0;
  ShmBlock* shm = nullptr;
  std::string shm_path;
};

namespace {

size_t env_size(const char* name, size_t dflt) {
  const char* s = getenv(name);
  if (!s || !*s) return dflt;
  char* end = nullptr;
  unsigned long long v = strtoull(s, &end, 10);
  return end == s ? dflt : static_cast<size_t>(v);
}

// Largest message (in wire bytes) for which one-shot beats two-shot, from the measured sweeps in profiles/
// (8xB200 NVSwitch): one-shot moves (W-1)x the payload per rank but needs one barrier instead of two, so the
// crossover falls quickly with W.  At W=2 both move the same bytes and one-shot wins until HBM traffic dominates.
size_t default_oneshot_max(int world) {
  if (world <= 2) return 16u << 20;
  if (world <= 4) return 2u << 20;
  return 512u << 10;
}

// Arena layout for a given world size; fills d.stage_off / d.slice_cap.
void layout(b2_comm* c, int world, size_t stage_bytes) {
  size_t cap = stage_bytes / (world + 1);
  cap &= ~static_cast<size_t>(255);
  c->stage_bytes = cap * (world + 1);
  const size_t stages0 = kLanes * kFlagRegionBytes;
  const int lanes = c->split_min_wire_bytes == ~static_cast<size_t>(0) ? 1 : kLanes;  // lane 1 costs memory only when enabled
  for (int lane = 0; lane < kLanes; ++lane) {
    CommDev& d = lane == 0 ? c->d : c->d1;
    d.slice_cap = cap;
    d.flag_off = lane * kFlagRegionBytes;
    d.stage_off[0] = stages0 + (2 * lane + 0) * c->stage_bytes;
    d.stage_off[1] = stages0 + (2 * lane + 1) * c->stage_bytes;
  }
  c->arena_bytes = stages0 + 2 * lanes * c->stage_bytes;
}

int alloc_rank_resources(b2_comm* c, int rank, int world, int device, size_t stage_bytes) {
  c->device = device;
  c->d.rank = c->d1.rank = rank;
  c->d.world = c->d1.world = world;
  if (stage_bytes == 0) stage_bytes = env_size("B2_STAGE_MB", kDefaultStageBytes >> 20) << 20;
  if (stage_bytes < (static_cast<size_t>(world + 1) << 12))
    return fail(B2_EINVAL, "stage_bytes=%zu too small for world=%d", stage_bytes, world);
  c->d.timeout_ns = c->d1.timeout_ns = env_size("B2_TIMEOUT_MS", kDefaultTimeoutNs / 1000000ull) * 1000000ull;
  // Measured on 4xB200 (profiles/r01_lane_split_w4.md): two symmetric half-collectives hit their barriers at the same time,
  // so the split buys nothing at 4-32 MiB and costs 6-19 % above 64 MiB.  Off by default; B2_SPLIT_MIN_BYTES enables it.
  c->split_min_wire_bytes = env_size("B2_SPLIT_MIN_BYTES", ~static_cast<size_t>(0));
  c->stagger = env_size("B2_STAGGER", 1) != 0;
  layout(c, world, stage_bytes);
  c->max_ctas = static_cast<int>(env_size("B2_MAX_CTAS", 0));
  c->oneshot_max_wire_bytes = env_size("B2_ONESHOT_MAX_BYTES", default_oneshot_max(world));
  c->auto_twoshot = env_size("B2_AUTO_TWOSHOT", B2_ALGO_TWOSHOT) == B2_ALGO_TWOSHOT_PULL ? B2_ALGO_TWOSHOT_PULL : B2_ALGO_TWOSHOT;
  B2_CUDA(cudaSetDevice(device));
  B2_CUDA(cudaMalloc(&c->arena, c->arena_bytes));
  B2_CUDA(cudaMemset(c->arena, 0, kLanes * kFlagRegionBytes));
  B2_CUDA(cudaMalloc(&c->counters, 256));
  B2_CUDA(cudaMemset(c->counters, 0, 256));
  c->d.opseq = c->counters;
  c->d.done = c->counters + 32;  // a different 128 B line
  c->d1.opseq = c->counters + 16;
  c->d1.done = c->counters + 48;
  c->d.stagger_ctr = c->d1.stagger_ctr = c->counters + 60;
  B2_CUDA(cudaStreamCreateWithFlags(&c->lane_stream, cudaStreamNonBlocking));
  B2_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  B2_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  B2_CUDA(cudaHostAlloc(&c->status_host, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(c->status_host, 0, 64);
  void* sdev = nullptr;
  B2_CUDA(cudaHostGetDevicePointer(&sdev, c->status_host, 0));
  c->d.status = c->d1.status = static_cast<uint32_t*>(sdev);
  B2_CUDA(cudaDeviceSynchronize());
  c->arena_of[rank] = static_cast<uint8_t*>(c->arena);
  return B2_OK;
}

void rotate_peers(b2_comm* c) {
  for (int jj = 0; jj < c->d.world; ++jj) {
    c->d.peer[jj] = c->d1.peer[jj] = c->arena_of[(c->d.rank + jj) % c->d.world];
    c->d.abs[jj] = c->d1.abs[jj] = c->arena_of[jj];
  }
}

void free_rank_resources(b2_comm* c) {
  if (c->device >= 0) cudaSetDevice(c->device);
  if (c->arena) cudaFree(c->arena);
  if (c->counters) cudaFree(c->counters);
  if (c->lane_stream) cudaStreamDestroy(c->lane_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  c->lane_stream = nullptr;
  c->ev_fork = c->ev_join = nullptr;
  if (c->trace_dev) cudaFree(c->trace_dev);
  c->trace_dev = nullptr;
  if (c->status_host) cudaFreeHost(c->status_host);
  c->arena = nullptr;
  c->counters = nullptr;
  c->status_host = nullptr;
}

int grid_for(const b2_comm* c, unsigned long long vecs_per_cta_dim, int unroll) {
  // Enough CTAs that each thread has work, capped so the collective leaves SMs to the backward
  // pass it overlaps with.  Deterministic in (n, world, max_ctas) => identical on every rank.
  // Default cap: 64 CTAs; 128 once 64 CTAs would each loop more than twice (from ~16 MiB fp32 buckets at W=8), where the
  // sweeps in profiles/r01_sweep_w*.jsonl show the 128-CTA grid 3-6 % faster.  One CTA per SM (512 threads, <= 128 regs).
  const int dflt = vecs_per_cta_dim > 64ull * kThreads * static_cast<unsigned long long>(unroll) * 2ull ? 128 : 64;
  const int cap = c->max_ctas > 0 ? (c->max_ctas > kMaxCtas ? kMaxCtas : c->max_ctas) : dflt;
  unsigned long long per = static_cast<unsigned long long>(kThreads) * unroll;
  unsigned long long g = (vecs_per_cta_dim + per - 1) / per;
  if (g < 1) g = 1;
  if (g > static_cast<unsigned long long>(cap)) g = cap;
  return static_cast<int>(g);
}

int unroll_for_world(int w) { return w >= 8 ? 1 : (w >= 4 ? 2 : (w >= 2 ? 4 : 8)); }

template <int MODE, int W>
cudaError_t launch_oneshot(const CommDev& d, int grid, void* buf, unsigned long long n, float scale,
                           cudaStream_t s) {
  k_oneshot<MODE, W><<<grid, kThreads, 0, s>>>(d, buf, n, scale);
  return cudaGetLastError();
}
template <int MODE, int W>
cudaError_t launch_twoshot(const CommDev& d, int grid, void* buf, unsigned long long n, float scale,
                           cudaStream_t s) {
  k_twoshot<MODE, W><<<grid, kThreads, 0, s>>>(d, buf, n, scale);
  return cudaGetLastError();
}

template <int MODE, int W>
cudaError_t launch_twoshot_pull(const CommDev& d, int grid, void* buf, unsigned long long n, float scale,
                                cudaStream_t s) {
  k_twoshot_pull<MODE, W><<<grid, kThreads, 0, s>>>(d, buf, n, scale);
  return cudaGetLastError();
}

// kind: 1 = one-shot, 2 = two-shot (push-scatter), 3 = two-shot (pull-pull)
template <int MODE>
cudaError_t launch_by_world(const CommDev& d, int kind, int grid, void* buf, unsigned long long n,
                            float scale, cudaStream_t s) {
#define B2_CASE(Wv)                                                                    \
  case Wv:                                                                             \
    return kind == 1   ? launch_oneshot<MODE, Wv>(d, grid, buf, n, scale, s)           \
           : kind == 2 ? launch_twoshot<MODE, Wv>(d, grid, buf, n, scale, s)           \
                       : launch_twoshot_pull<MODE, Wv>(d, grid, buf, n, scale, s);
  switch (d.world) {
    B2_CASE(2)
    B2_CASE(3)
    B2_CASE(4)
    B2_CASE(5)
    B2_CASE(6)
    B2_CASE(7)
    B2_CASE(8)
    default:
      return cudaErrorInvalidValue;
  }
#undef B2_CASE
}

template <int MODE>
cudaError_t launch_local(void* buf, unsigned long long n, float scale, cudaStream_t s) {
  // TMA-staged path for 16 B-aligned buckets.  Measured on B200 (profiles/r01_local_pass_tma_vs_plain.jsonl): the ring
  // needs several tiles per CTA to pay for its prologue - slower than the plain kernel below 25 MiB, equal to 168 MiB,
  // ahead from 512 MiB - so it is the default from 256 MiB up.  B2_LOCAL_TMA_MIN_MB moves the threshold (0 = always
  // for >= 1 MiB, e.g. for the parity tests), B2_LOCAL_TMA=0 disables it.
  static const bool use_tma = env_size("B2_LOCAL_TMA", 1) != 0;
  static const unsigned long long tma_min = env_size("B2_LOCAL_TMA_MIN_MB", 256) << 20;
  const unsigned long long nbytes = n * (MODE == B2_BF16 ? 2 : 4);
  if (use_tma && (reinterpret_cast<uintptr_t>(buf) & 15u) == 0 && nbytes >= (1ull << 20) && nbytes >= tma_min) {
    const unsigned long long ntiles = nbytes / tma::kTileBytes;
    unsigned long long g = ntiles < 148ull * 2 ? ntiles : 148ull * 2;  // persistent: 2 CTAs (2 x 64 KiB rings) per SM
    constexpr int kSmem = tma::kStages * tma::kTileBytes;
    static std::atomic<unsigned> configured{0};  // bit d: the 64 KiB opt-in has been set on device d
    int devno = 0;
    cudaGetDevice(&devno);
    if (!(configured.load(std::memory_order_relaxed) & (1u << (devno & 31)))) {
      const cudaError_t attr = cudaFuncSetAttribute(k_local_pass_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      if (attr != cudaSuccess) return attr;
      configured.fetch_or(1u << (devno & 31), std::memory_order_relaxed);
    }
    k_local_pass_tma<MODE><<<static_cast<int>(g), tma::kTmaThreads, kSmem, s>>>(buf, n, scale);
    return cudaGetLastError();
  }
  const unsigned long long V = (n + 7) / 8;
  unsigned long long g = (V + kThreads * 4ull - 1) / (kThreads * 4ull);
  if (g < 1) g = 1;
  if (g > 148ull * 4) g = 148ull * 4;  // 4 resident CTAs per SM keep ~64 KiB of loads in flight per SM
  k_local_pass<MODE><<<static_cast<int>(g), kThreads, 0, s>>>(buf, n, scale);
  return cudaGetLastError();
}

cudaError_t launch_mode(const CommDev& d, int mode, int kind, int grid, void* buf, unsigned long long n, float scale,
                        cudaStream_t s) {
  switch (mode) {
    case B2_F32_WIRE_BF16:
      return launch_by_world<B2_F32_WIRE_BF16>(d, kind, grid, buf, n, scale, s);
    case B2_F32:
      return launch_by_world<B2_F32>(d, kind, grid, buf, n, scale, s);
    default:
      return launch_by_world<B2_BF16>(d, kind, grid, buf, n, scale, s);
  }
}

size_t elem_bytes(int mode) { return mode == B2_BF16 ? 2 : 4; }
size_t wire_vec_bytes(int mode) { return mode == B2_F32 ? 32 : 16; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int b2_version(void) { return B2_ABI_VERSION; }

const char* b2_last_error(void) { return g_err.c_str(); }

int b2_comm_create_local(b2_comm_t** out, int world, const int* devices, size_t stage_bytes) {
  if (!out || !devices || world < 1 || world > B2_MAX_WORLD)
    return fail(B2_EINVAL, "b2_comm_create_local: bad arguments (world=%d)", world);
  int prev = -1;
  cudaGetDevice(&prev);
  b2_comm* cs[B2_MAX_WORLD] = {};
  int rc = B2_OK;
  for (int r = 0; r < world && rc == B2_OK; ++r) {
    cs[r] = new (std::nothrow) b2_comm();
    if (!cs[r]) {
      rc = fail(B2_ESYS, "out of host memory");
      break;
    }
    cs[r]->local_world = true;
    rc = alloc_rank_resources(cs[r], r, world, devices[r], stage_bytes);
  }
  for (int a = 0; a < world && rc == B2_OK; ++a) {
    for (int b = 0; b < world && rc == B2_OK; ++b) {
      if (devices[a] == devices[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
      if (!can) {
        rc = fail(B2_ENOPEER, "device %d cannot access device %d over P2P", devices[a], devices[b]);
        break;
      }
      cudaSetDevice(devices[a]);
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
      } else if (e != cudaSuccess) {
        rc = fail(B2_ECUDA, "cudaDeviceEnablePeerAccess(%d->%d): %s", devices[a], devices[b],
                  cudaGetErrorString(e));
      }
    }
  }
  if (rc == B2_OK) {
    for (int a = 0; a < world; ++a)
      for (int b = 0; b < world; ++b) cs[a]->arena_of[b] = static_cast<uint8_t*>(cs[b]->arena);
    for (int r = 0; r < world; ++r) {
      rotate_peers(cs[r]);
      out[r] = cs[r];
    }
  } else {
    for (int r = 0; r < world; ++r)
      if (cs[r]) {
        free_rank_resources(cs[r]);
        delete cs[r];
      }
  }
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

int b2_comm_create(b2_comm_t** out, int rank, int world, int device, const char* shm_name,
                   uint64_t epoch, size_t stage_bytes, int timeout_ms) {
  if (!out || world < 1 || world > B2_MAX_WORLD || rank < 0 || rank >= world || device < 0)
    return fail(B2_EINVAL, "b2_comm_create: bad arguments (rank=%d world=%d device=%d)", rank, world,
                device);
  if (world > 1 && (!shm_name || !*shm_name))
    return fail(B2_EINVAL, "b2_comm_create: shm_name is required when world > 1");
  const double deadline = now_s() + (timeout_ms > 0 ? timeout_ms : 120000) * 1e-3;
  int prev = -1;
  cudaGetDevice(&prev);
  b2_comm* c = new (std::nothrow) b2_comm();
  if (!c) return fail(B2_ESYS, "out of host memory");
  int rc = alloc_rank_resources(c, rank, world, device, stage_bytes);
  int fd = -1;
  if (rc == B2_OK && world > 1) {
    char path[256];
    snprintf(path, sizeof(path), "%s%s.e%llu", shm_name[0] == '/' ? "" : "/", shm_name,
             static_cast<unsigned long long>(epoch));
    c->shm_path = path;
    fd = shm_open(path, O_CREAT | O_RDWR, 0600);
    if (fd < 0) rc = fail(B2_ESYS, "shm_open(%s): %s", path, strerror(errno));
    if (rc == B2_OK && ftruncate(fd, sizeof(ShmBlock)) != 0)
      rc = fail(B2_ESYS, "ftruncate(%s): %s", path, strerror(errno));
    if (rc == B2_OK) {
      void* m = mmap(nullptr, sizeof(ShmBlock), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (m == MAP_FAILED)
        rc = fail(B2_ESYS, "mmap(%s): %s", path, strerror(errno));
      else
        c->shm = static_cast<ShmBlock*>(m);
    }
    if (fd >= 0) close(fd);
  }
  if (rc == B2_OK && world > 1) {
    ShmBlock* sb = c->shm;
    ShmSlot& me = sb->slot[rank];
    cudaError_t e = cudaIpcGetMemHandle(&me.handle, c->arena);
    if (e != cudaSuccess) rc = fail(B2_ECUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    if (rc == B2_OK) {
      me.device = device;
      memset(me.bus_id, 0, sizeof(me.bus_id));
      cudaDeviceGetPCIBusId(me.bus_id, sizeof(me.bus_id), device);
      me.pid = static_cast<int>(getpid());
      me.arena_bytes = c->arena_bytes;
      if (rank == 0) {
        sb->epoch = epoch;
        sb->world = world;
        sb->magic.store(kShmMagic, std::memory_order_release);
      }
      me.ready.store(1, std::memory_order_release);
    }
    // wait for every rank's slot
    for (int r = 0; r < world && rc == B2_OK; ++r) {
      while (sb->slot[r].ready.load(std::memory_order_acquire) != 1) {
        if (now_s() > deadline) {
          rc = fail(B2_ETIMEOUT, "rendezvous timed out waiting for rank %d on %s", r, c->shm_path.c_str());
          break;
        }
        usleep(200);
      }
    }
    // map every peer's arena
    for (int r = 0; r < world && rc == B2_OK; ++r) {
      if (r == rank) continue;
      const ShmSlot& ps = sb->slot[r];
      if (ps.arena_bytes != c->arena_bytes) {
        rc = fail(B2_EINVAL, "rank %d uses arena_bytes=%llu, this rank %zu (stage size must match)", r,
                  ps.arena_bytes, c->arena_bytes);
        break;
      }
      // Resolve the peer's GPU in THIS process's numbering by bus id.  If it is not visible here (each "node" of a
      // multi-node-on-one-box job gets its own CUDA_VISIBLE_DEVICES) the P2P query is impossible, but CUDA >= 10.1 can
      // still open an IPC mapping of an invisible peer's memory, so we just try.
      int peer_local = -1;
      if (cudaDeviceGetByPCIBusId(&peer_local, ps.bus_id) != cudaSuccess) {
        cudaGetLastError();
        peer_local = -1;
      }
      if (peer_local >= 0 && peer_local != device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, device, peer_local);
        if (!can) {
          rc = fail(B2_ENOPEER, "device %d cannot access rank %d's device %s over P2P", device, r, ps.bus_id);
          break;
        }
      }
      void* p = nullptr;
      e = cudaIpcOpenMemHandle(&p, ps.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        rc = fail(B2_ECUDA, "cudaIpcOpenMemHandle(rank %d, device %d): %s", r, ps.device,
                  cudaGetErrorString(e));
        break;
      }
      c->arena_of[r] = static_cast<uint8_t*>(p);
      c->peer_is_ipc[r] = true;
    }
    if (rc == B2_OK) {
      sb->mapped.fetch_add(1, std::memory_order_acq_rel);
      while (sb->mapped.load(std::memory_order_acquire) < world) {
        if (now_s() > deadline) {
          rc = fail(B2_ETIMEOUT, "rendezvous timed out waiting for peers to map (%d/%d)",
                    sb->mapped.load(), world);
          break;
        }
        usleep(200);
      }
    }
    // everyone holds a mapping now: the name can go (the memory lives until the last munmap)
    if (rc == B2_OK && rank == 0) shm_unlink(c->shm_path.c_str());
  }
  if (rc != B2_OK) {
    std::string keep = g_err;
    for (int r = 0; r < world; ++r)
      if (c->peer_is_ipc[r]) cudaIpcCloseMemHandle(c->arena_of[r]);
    if (c->shm) munmap(c->shm, sizeof(ShmBlock));
    if (rank == 0 && !c->shm_path.empty()) shm_unlink(c->shm_path.c_str());
    free_rank_resources(c);
    delete c;
    g_err = keep;
  } else {
    rotate_peers(c);
    *out = c;
  }
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

int b2_comm_destroy(b2_comm_t* c) {
  if (!c) return B2_OK;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->shm) {
    // nobody frees its arena while a peer may still have kernels reading it
    ShmBlock* sb = c->shm;
    sb->departed.fetch_add(1, std::memory_order_acq_rel);
    const double deadline = now_s() + 10.0;
    while (sb->departed.load(std::memory_order_acquire) < c->d.world && now_s() < deadline) usleep(200);
    for (int r = 0; r < c->d.world; ++r)
      if (c->peer_is_ipc[r]) cudaIpcCloseMemHandle(c->arena_of[r]);
    munmap(c->shm, sizeof(ShmBlock));
  }
  free_rank_resources(c);
  delete c;
  if (prev >= 0) cudaSetDevice(prev);
  return B2_OK;
}

int b2_comm_rank(const b2_comm_t* c) { return c ? c->d.rank : B2_EINVAL; }
int b2_comm_world(const b2_comm_t* c) { return c ? c->d.world : B2_EINVAL; }
int b2_comm_device(const b2_comm_t* c) { return c ? c->device : B2_EINVAL; }

int b2_comm_set_timeout_ms(b2_comm_t* c, int timeout_ms) {
  if (!c || timeout_ms <= 0) return fail(B2_EINVAL, "b2_comm_set_timeout_ms: bad arguments");
  c->d.timeout_ns = c->d1.timeout_ns = static_cast<unsigned long long>(timeout_ms) * 1000000ull;
  return B2_OK;
}

int b2_comm_set_max_ctas(b2_comm_t* c, int max_ctas) {
  if (!c || max_ctas < 0) return fail(B2_EINVAL, "b2_comm_set_max_ctas: bad arguments");
  c->max_ctas = max_ctas;
  return B2_OK;
}

int b2_comm_status(const b2_comm_t* c) {
  if (!c) return fail(B2_EINVAL, "null communicator");
  const uint32_t s = *reinterpret_cast<volatile uint32_t*>(c->status_host);
  if (s == 0) return B2_OK;
  return fail(-static_cast<int>(s), "rank %d: a kernel gave up waiting for a peer (code %d)", c->d.rank,
              -static_cast<int>(s));
}

uint64_t b2_comm_launch_count(const b2_comm_t* c) { return c ? c->launches : 0; }

int b2_comm_trace(b2_comm_t* c, int enable, uint64_t* out, int max_ctas) {
  if (!c) return fail(B2_EINVAL, "null communicator");
  DeviceGuard g(c->device);
  if (enable && !c->trace_dev) B2_CUDA(cudaMalloc(&c->trace_dev, sizeof(unsigned long long) * kMaxCtas * 8));
  if (out && max_ctas > 0 && c->trace_dev) {
    const int n = max_ctas < kMaxCtas ? max_ctas : kMaxCtas;
    B2_CUDA(cudaMemcpy(out, c->trace_dev, sizeof(unsigned long long) * n * 8, cudaMemcpyDeviceToHost));
  }
  if (enable) B2_CUDA(cudaMemset(c->trace_dev, 0, sizeof(unsigned long long) * kMaxCtas * 8));  // no stale CTAs
  c->d.trace = enable ? c->trace_dev : nullptr;
  return B2_OK;
}

int b2_local_pass(void* buf, size_t n_elems, int mode, float scale, int device, void* stream) {
  if (n_elems == 0) return B2_OK;
  if (!buf) return fail(B2_EINVAL, "b2_local_pass: null buffer");
  DeviceGuard g(device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (mode) {
    case B2_F32_WIRE_BF16:
      e = launch_local<B2_F32_WIRE_BF16>(buf, n_elems, scale, s);
      break;
    case B2_F32:
      e = launch_local<B2_F32>(buf, n_elems, scale, s);
      break;
    case B2_BF16:
      e = launch_local<B2_BF16>(buf, n_elems, scale, s);
      break;
    default:
      return fail(B2_EINVAL, "unknown mode %d", mode);
  }
  if (e != cudaSuccess) return fail(B2_ECUDA, "k_local_pass launch: %s", cudaGetErrorString(e));
  return B2_OK;
}

int b2_allreduce(b2_comm_t* c, void* buf, size_t n_elems, int mode, float scale, int algo, void* stream) {
  if (!c) return fail(B2_EINVAL, "null communicator");
  if (mode != B2_F32_WIRE_BF16 && mode != B2_F32 && mode != B2_BF16)
    return fail(B2_EINVAL, "unknown mode %d", mode);
  if (algo != B2_ALGO_AUTO && algo != B2_ALGO_ONESHOT && algo != B2_ALGO_TWOSHOT && algo != B2_ALGO_TWOSHOT_PULL)
    return fail(B2_EINVAL, "unknown algo %d", algo);
  if (n_elems == 0) return B2_OK;
  if (!buf) return fail(B2_EINVAL, "b2_allreduce: null buffer");
  if (*reinterpret_cast<volatile uint32_t*>(c->status_host) != 0)
    return fail(B2_ESTATE, "communicator poisoned by an earlier peer-wait timeout");
  const int W = c->d.world;
  if (W == 1) {
    if (mode == B2_F32 && scale == 1.0f) return B2_OK;  // identity
    int rc = b2_local_pass(buf, n_elems, mode, scale, c->device, stream);
    if (rc == B2_OK) c->launches++;
    return rc;
  }
  DeviceGuard g(c->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t wvb = wire_vec_bytes(mode);
  const size_t cap_vecs = c->d.slice_cap / wvb;  // vecs one region can hold
  const int U = unroll_for_world(W);
  uint8_t* p = static_cast<uint8_t*>(buf);
  size_t left = n_elems;
  while (left > 0) {
    const unsigned long long V_left = (left + 7) / 8;
    int kind;
    if (algo == B2_ALGO_AUTO) {
      kind = (V_left * wvb <= c->oneshot_max_wire_bytes && V_left <= cap_vecs) ? B2_ALGO_ONESHOT : c->auto_twoshot;
    } else {
      kind = algo;
    }
    const bool oneshot = kind == B2_ALGO_ONESHOT;
    const unsigned long long max_vecs = oneshot ? cap_vecs : cap_vecs * W;
    const unsigned long long V = V_left < max_vecs ? V_left : max_vecs;
    const size_t n = V == V_left ? left : static_cast<size_t>(V) * 8;
    cudaError_t e;
    if (!oneshot && V * wvb >= c->split_min_wire_bytes && V >= 2ull * W) {
      // two concurrent half-collectives: lane 1 (second half) on the internal stream, lane 0 on the caller's
      const unsigned long long V0 = (V / 2 + W - 1) / W * W;  // whole vecs, W-aligned so both halves slice evenly
      const size_t n0 = static_cast<size_t>(V0) * 8;
      B2_CUDA(cudaEventRecord(c->ev_fork, s));
      B2_CUDA(cudaStreamWaitEvent(c->lane_stream, c->ev_fork, 0));
      // both halves must be co-resident (1 CTA per SM, 148 SMs): at most 64 CTAs each; lane 0 first so it is resident
      // before lane 1's CTAs start waiting for its scatter (push variant only: the stagger hooks live in k_twoshot)
      const bool stagger = c->stagger && kind == B2_ALGO_TWOSHOT;
      CommDev d0 = c->d, d1 = c->d1;
      d0.stagger_role = stagger ? 1 : 0;
      d1.stagger_role = stagger ? 2 : 0;
      int g0 = grid_for(c, V0 / W, U), g1 = grid_for(c, (V - V0 + W - 1) / W, U);
      if (g0 > 64) g0 = 64;
      if (g1 > 64) g1 = 64;
      e = launch_mode(d0, mode, kind, g0, p, n0, scale, s);
      if (e == cudaSuccess) e = launch_mode(d1, mode, kind, g1, p + n0 * elem_bytes(mode), n - n0, scale, c->lane_stream);
      if (e != cudaSuccess) return fail(B2_ECUDA, "allreduce kernel launch: %s", cudaGetErrorString(e));
      B2_CUDA(cudaEventRecord(c->ev_join, c->lane_stream));
      B2_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
      c->launches += 2;
    } else {
      const unsigned long long per_cta_dim = oneshot ? V : (V + W - 1) / W;
      e = launch_mode(c->d, mode, kind, grid_for(c, per_cta_dim, U), p, n, scale, s);
      if (e != cudaSuccess) return fail(B2_ECUDA, "allreduce kernel launch: %s", cudaGetErrorString(e));
      c->launches++;
    }
    p += n * elem_bytes(mode);
    left -= n;
  }
  return B2_OK;
}

int b2_broadcast(b2_comm_t* c, void* buf, size_t bytes, int root, void* stream) {
  if (!c) return fail(B2_EINVAL, "null communicator");
  if (root < 0 || root >= c->d.world) return fail(B2_EINVAL, "b2_broadcast: root %d out of range", root);
  if (bytes == 0 || c->d.world == 1) return B2_OK;
  if (!buf) return fail(B2_EINVAL, "b2_broadcast: null buffer");
  if (*reinterpret_cast<volatile uint32_t*>(c->status_host) != 0)
    return fail(B2_ESTATE, "communicator poisoned by an earlier peer-wait timeout");
  DeviceGuard g(c->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t cap = c->stage_bytes & ~static_cast<size_t>(15);
  uint8_t* p = static_cast<uint8_t*>(buf);
  size_t left = bytes;
  while (left > 0) {
    const size_t n = left < cap ? left : cap;
    const int grid = grid_for(c, (n + 15) / 16, 1);
    k_broadcast<<<grid, kThreads, 0, s>>>(c->d, p, n, root);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2_ECUDA, "broadcast kernel launch: %s", cudaGetErrorString(e));
    c->launches++;
    p += n;
    left -= n;
  }
  return B2_OK;
}

int b2_barrier(b2_comm_t* c, void* stream) {
  if (!c) return fail(B2_EINVAL, "null communicator");
  if (c->d.world == 1) return B2_OK;
  if (*reinterpret_cast<volatile uint32_t*>(c->status_host) != 0)
    return fail(B2_ESTATE, "communicator poisoned by an earlier peer-wait timeout");
  DeviceGuard g(c->device);
  k_barrier<<<1, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(c->d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(B2_ECUDA, "barrier kernel launch: %s", cudaGetErrorString(e));
  c->launches++;
  return B2_OK;
}

}  // extern "C"
