// shim.h — TEST INFRASTRUCTURE. Lets g++ compile the CUDA C source that the product generates for the compiled
// resident executor (csrc/ew_jit.cpp tape_jit; dumped with CG_TAPE_JIT_DUMP in plan-only mode) as ordinary C++, so that
// tests/test_tape_emulator.py can EXECUTE the generated code on the CPU — 512 OS threads standing in for the CTA's
// threads, std::barrier for __syncthreads(), one global array for the dynamic shared memory — and compare it with the
// oracle. Every rounding intrinsic maps to the single IEEE operation it names (build with -ffp-contract=off); expf and
// tanhf are the host libm's, which is also what the oracle calls. Nothing in the product includes this file.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__
#define __launch_bounds__(...)

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct EmuThreadIdx { unsigned x; };
extern thread_local EmuThreadIdx threadIdx;
extern std::barrier<>* emu_barrier;
static inline void __syncthreads() { emu_barrier->arrive_and_wait(); }

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline float __int_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
