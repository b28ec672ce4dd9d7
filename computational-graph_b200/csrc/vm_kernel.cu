// vm_kernel.cu — the resident tape executor for latency-bound graphs (BASELINE config 3: the README LSTM, d <= 30).
//
// A graph whose matrices hold a few hundred elements is bound by launch latency, not by any roofline: the README LSTM
// in exact mode is 103 kernels per iteration at ~1.7 us of CUDA-graph node latency each, for ~10 us of arithmetic.
// Here ONE CTA stays resident for the whole `minimize` call and walks the compiled tape itself: every step (a fused
// elementwise program, a strict product, a scalar<-matrix mean) is executed by the CTA's threads, steps are separated
// by __syncthreads(), and the iteration loop of optimize.rs:64-71 runs inside the kernel. The working set (< 100 KB)
// lives in this SM's L1 / the L2; one SM's L1 is coherent with its own stores, which is all a single CTA needs.
//
// Arithmetic is the per-step kernels' arithmetic, instruction for instruction (ew_run is the same interpreter, the
// product is the k-ascending separate-rounding loop of gemm_strict_kernel = cpu/ops.cpp:103-114, the mean is the
// sequential sum of cpu/ops.cpp:138-145), so a plan gives the same bits on either executor (tests/test_gpu_graph.py).
#include "ew_interp.cuh"
#include "tape.h"

namespace cg {

namespace {

constexpr int VM_THREADS = 512;  // <= 128 registers per thread for the interpreter's 12 x float4 register file

__device__ __forceinline__ void vm_copy(float* dst, const float* src, unsigned n) {
  for (unsigned i = threadIdx.x; i < n; i += VM_THREADS) dst[i] = src[i];
}

template <bool TA, bool TB>
__device__ __forceinline__ void vm_gemm(const VmStep& st) {
  const int total = st.m * st.n;
  for (int o = threadIdx.x; o < total; o += VM_THREADS) {
    const int i = o % st.m, j = o / st.m;
    const float* a = TA ? st.a + (size_t)i * st.lda : st.a + i;
    const float* b = TB ? st.b + j : st.b + (size_t)j * st.ldb;
    const int sa = TA ? 1 : st.lda, sb = TB ? st.ldb : 1;
    float acc = 0.0f;
#pragma unroll 4
    for (int kk = 0; kk < st.k; kk++) acc = __fadd_rn(acc, __fmul_rn(a[(size_t)kk * sa], b[(size_t)kk * sb]));
    float* c = st.c + (size_t)j * st.ldc + i;
    if (st.epi == 1) *c = __fsub_rn(*c, __fmul_rn(acc, st.lr));  // optimize.rs:149-152, two roundings
    else if (st.epi == 2) *c = __fadd_rn(*c, acc);
    else *c = acc;
  }
}

__global__ void __launch_bounds__(VM_THREADS, 1)
    tape_vm_kernel(const VmStep* steps, const EwProgram* progs, int nsteps, int iters, const float* root, unsigned root_n,
                   int root_is_leaf, float* trace) {
  __shared__ float red[kStrictReduceMax];
  for (int it = 0; it < iters; it++) {
    // the traced value of an iteration is its forward value, before the update (Q12): a leaf root is read first
    if (trace && root_is_leaf) {
      vm_copy(trace + (size_t)it * root_n, root, root_n);
      __syncthreads();
    }
    for (int s = 0; s < nsteps; s++) {
      const VmStep& st = steps[s];
      const int kind = st.kind;
      if (kind == VmStep::EW) {
        const EwProgram& p = progs[st.prog];
        const size_t n = p.n;
        for (size_t base = (size_t)threadIdx.x * 4; base < n; base += (size_t)VM_THREADS * 4) ew_run<4>(p, base, 0);
      } else if (kind == VmStep::GEMM) {
        if (st.ta) {
          if (st.tb) vm_gemm<true, true>(st);
          else vm_gemm<true, false>(st);
        } else {
          if (st.tb) vm_gemm<false, true>(st);
          else vm_gemm<false, false>(st);
        }
      } else {
        // scalar <- matrix mean: stage through shared memory, then the reference's sequential sum
        const unsigned n = (unsigned)st.n;
        for (unsigned i = threadIdx.x; i < n; i += VM_THREADS) red[i] = st.a[i];
        __syncthreads();
        if (threadIdx.x == 0) {
          float sum = 0.0f;
          for (unsigned i = 0; i < n; i++) sum = __fadd_rn(sum, red[i]);
          st.c[0] = __fdiv_rn(sum, (float)n);
        }
      }
      __syncthreads();
    }
    if (trace && !root_is_leaf) {
      vm_copy(trace + (size_t)it * root_n, root, root_n);
      __syncthreads();
    }
  }
}

}  // namespace

void launch_tape_vm(const VmStep* steps, const EwProgram* progs, int nsteps, int iters, const float* root, size_t root_n,
                    bool root_is_leaf, float* trace, cudaStream_t stream) {
  if (nsteps <= 0 || iters <= 0) return;
  tape_vm_kernel<<<1, VM_THREADS, 0, stream>>>(steps, progs, nsteps, iters, root, (unsigned)root_n, root_is_leaf ? 1 : 0, trace);
  CG_CUDA(cudaGetLastError());
}

}  // namespace cg
