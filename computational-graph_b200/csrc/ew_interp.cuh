// ew_interp.cuh — the fused-elementwise bytecode interpreter (device side).
//
// One thread evaluates a whole elementwise sub-tree (forward chain, its derivatives and the SGD update)
// for VEC consecutive-by-float4 elements, entirely in registers: the reference's one-kernel-per-node
// evaluation (src/gpu/ops.cu:11-80, 16-thread blocks) becomes one HBM pass per sub-tree.
//
// Instruction encoding: (opcode << 4) | operand, operand < 16 (register, immediate or output slot).
// The register file is a set of real registers: every (opcode, register) pair is its own switch case,
// so ptxas emits one jump table and no local-memory indexing.
#pragma once
#include "cg_common.h"

namespace cg {

#define CG_FOR_EACH_REG(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11)
static_assert(EW_REGS == 12, "CG_FOR_EACH_REG must list EW_REGS registers");

__device__ __forceinline__ float ew_sigmoid(float x) {
  // cpu/ops.cpp:72 — 1.0f / (1.0f + expf(-x)); IEEE division (no -use_fast_math), CUDA expf (<= 2 ulp)
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
__device__ __forceinline__ float ew_sgn_scalar(float x) {
  // f32::signum: 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives, NaN for NaN (ops.rs:332)
  if (x != x) return x;
  return (__float_as_uint(x) >> 31) ? -1.0f : 1.0f;
}

template <int VEC>
__device__ __forceinline__ void ew_load(const float* __restrict__ p, size_t base, size_t stride, size_t n,
                                        float (&dst)[VEC]) {
#pragma unroll
  for (int v = 0; v < VEC / 4; v++) {
    size_t b = base + (size_t)v * stride;
    if (b + 4 <= n) {
      float4 x = *reinterpret_cast<const float4*>(p + b);
      dst[v * 4 + 0] = x.x; dst[v * 4 + 1] = x.y; dst[v * 4 + 2] = x.z; dst[v * 4 + 3] = x.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) dst[v * 4 + j] = (b + j < n) ? p[b + j] : 0.0f;
    }
  }
}

template <int VEC>
__device__ __forceinline__ void ew_store(float* __restrict__ p, size_t base, size_t stride, size_t n,
                                         const float (&src)[VEC]) {
#pragma unroll
  for (int v = 0; v < VEC / 4; v++) {
    size_t b = base + (size_t)v * stride;
    if (b + 4 <= n) {
      *reinterpret_cast<float4*>(p + b) = make_float4(src[v * 4 + 0], src[v * 4 + 1], src[v * 4 + 2], src[v * 4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (b + j < n) p[b + j] = src[v * 4 + j];
    }
  }
}

// Runs program `p` for the VEC elements {base + v*stride + j : v < VEC/4, j < 4} (those below p.n).
// `Prog` is EwProgram in the constant bank (kernel parameter) or in global memory (tape executor).
template <int VEC, typename Prog>
__device__ __forceinline__ void ew_run(const Prog& p, size_t base, size_t stride) {
  float r[EW_REGS][VEC];
  float acc[VEC];
  const size_t n = p.n;
  const uint32_t n_in = p.n_in;
  const uint32_t smask = p.in_scalar_mask;

  // prefetch: every input load is issued before the first dependent instruction
#pragma unroll
  for (int k = 0; k < EW_MAX_IN; k++) {
    if (k < (int)n_in) {
      if ((smask >> k) & 1u) {
        float s = *p.in[k];
#pragma unroll
        for (int j = 0; j < VEC; j++) r[k][j] = s;
      } else {
        ew_load<VEC>(p.in[k], base, stride, n, r[k]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; j++) acc[j] = 0.0f;

  // Two-level dispatch, both levels dense => ptxas emits jump tables (BRX), not compare trees.
  const uint32_t n_ins = p.n_ins;
  for (uint32_t pc = 0; pc < n_ins; pc++) {
    const uint32_t ins = p.code[pc];
    const uint32_t k = ins & 15u;
    switch (ins >> 4) {
#define CG_REG_SWITCH(STMT)                                         \
  switch (k) {                                                      \
    case 0: { constexpr int K = 0; STMT } break;                    \
    case 1: { constexpr int K = 1; STMT } break;                    \
    case 2: { constexpr int K = 2; STMT } break;                    \
    case 3: { constexpr int K = 3; STMT } break;                    \
    case 4: { constexpr int K = 4; STMT } break;                    \
    case 5: { constexpr int K = 5; STMT } break;                    \
    case 6: { constexpr int K = 6; STMT } break;                    \
    case 7: { constexpr int K = 7; STMT } break;                    \
    case 8: { constexpr int K = 8; STMT } break;                    \
    case 9: { constexpr int K = 9; STMT } break;                    \
    case 10: { constexpr int K = 10; STMT } break;                  \
    default: { constexpr int K = 11; STMT } break;                  \
  }
      case OP_LDR:
        CG_REG_SWITCH(_Pragma("unroll") for (int j = 0; j < VEC; j++) acc[j] = r[K][j];)
        break;
      case OP_STR:
        CG_REG_SWITCH(_Pragma("unroll") for (int j = 0; j < VEC; j++) r[K][j] = acc[j];)
        break;
      case OP_ADD:
        CG_REG_SWITCH(_Pragma("unroll") for (int j = 0; j < VEC; j++) acc[j] = __fadd_rn(acc[j], r[K][j]);)
        break;
      case OP_SUB:
        CG_REG_SWITCH(_Pragma("unroll") for (int j = 0; j < VEC; j++) acc[j] = __fsub_rn(acc[j], r[K][j]);)
        break;
      case OP_RSUB:
        CG_REG_SWITCH(_Pragma("unroll") for (int j = 0; j < VEC; j++) acc[j] = __fsub_rn(r[K][j], acc[j]);)
        break;
      case OP_MUL:
        CG_REG_SWITCH(_Pragma("unroll") for (int j = 0; j < VEC; j++) acc[j] = __fmul_rn(acc[j], r[K][j]);)
        break;
#undef CG_REG_SWITCH
      case OP_LDI: {
        const float s = p.imm[k];
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = s;
      } break;
      case OP_MULI: {
        const float s = p.imm[k];
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = __fmul_rn(acc[j], s);
      } break;
      case OP_NEG:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = -acc[j];
        break;
      case OP_SQ:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = __fmul_rn(acc[j], acc[j]);
        break;
      case OP_ABS_M:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = acc[j] < 0.0f ? -acc[j] : acc[j];
        break;
      case OP_SGN_M:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = acc[j] < 0.0f ? -1.0f : 1.0f;
        break;
      case OP_ABS_S:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = fabsf(acc[j]);
        break;
      case OP_SGN_S:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = ew_sgn_scalar(acc[j]);
        break;
      case OP_SIGMOID:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = ew_sigmoid(acc[j]);
        break;
      case OP_TANH:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = tanhf(acc[j]);
        break;
      case OP_ONE_MINUS:
#pragma unroll
        for (int j = 0; j < VEC; j++) acc[j] = __fsub_rn(1.0f, acc[j]);
        break;
      case OP_STG:
        ew_store<VEC>(p.out[k], base, stride, n, acc);
        break;
      default: break;
    }
  }
}

}  // namespace cg
