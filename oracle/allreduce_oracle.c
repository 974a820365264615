/*
 * allreduce_oracle.c — CPU restatement of the data-parallel hot path.  TEST INFRASTRUCTURE ONLY:
 * nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object; the product path (torchx_b200/) never does.
 *
 * The arithmetic of this path is not in meta-pytorch/torchx itself (it has no native code, SURVEY.md
 * §0): `dist.ddp` launches workers (torchx/components/dist.py:261-308) whose gradients are averaged by
 * third-party torch 2.10/2.11 + NCCL 2.27/2.28 (uv.lock:4600-4601,2743-2744).  What is restated here:
 *
 *   b2o_compress / b2o_allreduce, mode B2O_F32_WIRE_BF16
 *       torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:57-93 (`_compress_hook`):
 *         compressed = buffer.to(bf16).div_(world_size)  ->  all_reduce(SUM)  ->  buffer.copy_(compressed)
 *       `div_` by a Python scalar on CUDA multiplies by the fp32 reciprocal
 *       (aten/src/ATen/native/cuda/BinaryDivTrueKernel.cu); for W = 2^k both forms are exact.
 *   mode B2O_F32
 *       torch/csrc/distributed/c10d/reducer.cpp `mark_variable_ready_dense`: with no comm hook the bucket
 *       view is filled by `mul_out(bucket_view, grad, 1/div_factor)` and then SUM-allreduced in fp32
 *       (torch/include/torch/csrc/distributed/c10d/default_comm_hooks.hpp:36-51).
 *   mode B2O_BF16
 *       `dist.all_reduce` on a bf16 tensor after a pre-scale (torch/distributed/.../default_hooks.py:18-33,
 *       `_allreduce_fut`: `tensor.div_(W)` then all_reduce), fp32 accumulate.
 *
 * Reduction ORDER is the one thing the backend leaves undefined (NCCL ring/tree/NVLS).  This oracle - and
 * the CUDA path - fix it: fp32 accumulation in rank order 0..W-1 starting from rank 0's value, one final
 * rounding.  At W = 2 that is bit-identical to the reference on any backend (one add, one rounding; pinned
 * by tests/golden/ddp_w2_*.npz which were produced by the reference launcher + stock DDP over gloo, see
 * tests/golden/make_golden.py).  At W >= 4 the backend's own order differs from run to run of topology, so
 * the pin there is a tolerance (tests/golden/ddp_w4_*.npz).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define B2O_F32_WIRE_BF16 0
#define B2O_F32 1
#define B2O_BF16 2

static inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* fp32 -> bf16, round to nearest even; NaN -> 0x7fff (what `cvt.rn.bf16.f32` produces; torch's CPU
 * conversion gives 0x7fc0 - tests compare NaNs as NaNs, not by payload). */
uint16_t b2o_bf16_rne(float f) {
  uint32_t u = f2u(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

float b2o_bf16_to_f32(uint16_t h) { return u2f(((uint32_t)h) << 16); }

/* wire(scale * x): what one rank contributes.  `x_bits` is the raw element (fp32 bits, or bf16 bits in
 * the low half for B2O_BF16).  Returned as the fp32 value of the wire element. */
static inline float compress1(int mode, uint32_t x_bits, float scale) {
  if (mode == B2O_F32) return u2f(x_bits) * scale;
  float a;
  if (mode == B2O_F32_WIRE_BF16)
    a = b2o_bf16_to_f32(b2o_bf16_rne(u2f(x_bits))); /* the .to(bf16) cast */
  else
    a = b2o_bf16_to_f32((uint16_t)x_bits);
  return b2o_bf16_to_f32(b2o_bf16_rne(a * scale)); /* the div_ result, rounded to bf16 */
}

void b2o_compress(int mode, const void* in, size_t n, float scale, float* out_f32) {
  for (size_t i = 0; i < n; ++i) {
    uint32_t bits = mode == B2O_BF16 ? ((const uint16_t*)in)[i] : ((const uint32_t*)in)[i];
    out_f32[i] = compress1(mode, bits, scale);
  }
}

/*
 * in[r] points at rank r's n-element bucket (fp32, or bf16 for B2O_BF16); out receives the value every
 * rank ends up with, in the bucket's dtype.  Returns 0, or -1 on a bad mode.
 */
int b2o_allreduce(int mode, int world, const void* const* in, size_t n, float scale, void* out) {
  if (mode < 0 || mode > 2 || world < 1) return -1;
  for (size_t i = 0; i < n; ++i) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) {
      uint32_t bits = mode == B2O_BF16 ? ((const uint16_t*)in[r])[i] : ((const uint32_t*)in[r])[i];
      float c = compress1(mode, bits, scale);
      s = r == 0 ? c : s + c; /* start from rank 0's value: keeps -0.0 and NaN payload order */
    }
    if (mode == B2O_F32) {
      ((float*)out)[i] = s;
    } else if (mode == B2O_F32_WIRE_BF16) {
      ((float*)out)[i] = b2o_bf16_to_f32(b2o_bf16_rne(s));
    } else {
      ((uint16_t*)out)[i] = b2o_bf16_rne(s);
    }
  }
  return 0;
}

/* bit-exact broadcast restatement (trivial; kept so every C-ABI entry point has an oracle twin) */
void b2o_broadcast(const void* root_buf, size_t bytes, void* out) { memmove(out, root_buf, bytes); }
