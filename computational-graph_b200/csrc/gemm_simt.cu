// gemm_simt.cu — FP32 SIMT matrix product C = op(A) . op(B), column-major, NN / NT / TN / TT.
//
// Replaces the reference's cublasSgemm call (src/gpu/ops.cu:82-141) and restates the CPU gemm
// (src/cpu/ops.cpp:88-116) for shapes the reference itself cannot run (non-square operands, Q7).
// Forward products are NN, the two backward products of optimize.rs:196-200 are NT (E . W^T) and
// TN (X^T . E). Two kernels:
//   * gemm_strict_kernel — one thread per output element, k ascending, separate multiply and add
//     roundings: bit-identical to the reference's naive triple loop. Used for small products where
//     the latency of the launch dominates anyway (README LSTM, d <= 64).
//   * gemm_tiled_kernel  — 128x128x16 shared-memory tiles in a 3-stage ring (cp.async where the operand is contiguous
//     along the tile, transposing register staging where it is not), 8x8 register micro-tiles,
//     FMA accumulation (k ascending per output element). FMA-pipe bound. The micro-tile is accumulated with the
//     packed FFMA2 of sm_100 (`__ffma2_rn`: two IEEE fma.rn per instruction, the A element broadcast from a scalar
//     register): 32 FFMA2 + 4 LDS.128 per k instead of 64 FFMA + 4 LDS.128, so the issue slots no longer cap the FMA
//     pipe (ncu, scalar version: issue active 77 %, 85 % of the issued instructions FFMA => pipe 66 % busy).
// The tensor-core TF32 path lives in gemm_tcgen05.cu.
#include "cg_common.h"

namespace cg {

__device__ __forceinline__ float epilogue_apply(const GemmEpilogue& epi, float acc, float old) {
  if (epi.mode == 1) {
    // optimize.rs:149-152: error *= lr; value -= error — two roundings, never an FMA
    return __fsub_rn(old, __fmul_rn(acc, epi.lr));
  }
  if (epi.mode == 2) return __fadd_rn(old, acc);
  return acc;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_strict_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ C, int m, int n, int k, int lda, int ldb,
                                                          int ldc, GemmEpilogue epi) {
  const int i = blockIdx.x * 16 + (threadIdx.x & 15);
  const int j = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= m || j >= n) return;
  float acc = 0.0f;
  for (int kk = 0; kk < k; kk++) {
    const float a = TA ? A[(size_t)i * lda + kk] : A[(size_t)kk * lda + i];
    const float b = TB ? B[(size_t)kk * ldb + j] : B[(size_t)j * ldb + kk];
    acc = __fadd_rn(acc, __fmul_rn(a, b));
  }
  float* c = C + (size_t)j * ldc + i;
  *c = epilogue_apply(epi, acc, epi.mode != 0 ? *c : 0.0f);
}

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, STAGES = 3;
constexpr size_t TILED_SMEM = (size_t)STAGES * 2 * BK * (BM + PAD) * sizeof(float);  // 50.7 KB: two CTAs per SM
static_assert(BM == BN, "the staging helpers assume equal tile widths");
using TileRow = float[BM + PAD];

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One (BK x 128) operand tile: element (kk, x) is op(M)(x0 + x, k0 + kk) for A (x = row of C) or op(M)(k0 + kk, x0 + x) for
// B (x = column of C), and lands in shared memory as S[kk][x].
//   MN_CONTIG (A not transposed, B transposed): the stored matrix is contiguous along x, so a whole, aligned tile goes
//     global -> shared with cp.async (16 bytes per copy, no registers, two copies per thread) and is in flight for the
//     whole of the next tile's arithmetic;
//   otherwise (contiguous along k) the tile is transposed on its way: float4 loads along k into registers at the top of
//     an iteration, scalar stores S[k][x] after the arithmetic. Edge tiles of either kind take the register path with
//     bounds checks.
template <bool MN_CONTIG>
__device__ __forceinline__ void tile_async(const float* __restrict__ M, int ld, int x0, int k0, TileRow* S) {
  static_assert(MN_CONTIG, "cp.async staging needs the tile contiguous along x");
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int f = threadIdx.x + s * 256, kk = f >> 5, x4 = (f & 31) * 4;
    cp_async16(&S[kk][x4], M + (size_t)(k0 + kk) * ld + x0 + x4);
  }
}
template <bool MN_CONTIG>
__device__ __forceinline__ void load_tile(const float* __restrict__ M, int ld, int x0, int k0, int xmax, int kmax,
                                          bool fast, float (&reg)[8]) {
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int f = threadIdx.x + s * 256;
    if (MN_CONTIG) {
      const int kk = f >> 5, x4 = (f & 31) * 4;
      const float* p = M + (size_t)(k0 + kk) * ld + x0 + x4;
#pragma unroll
      for (int c = 0; c < 4; c++) reg[s * 4 + c] = (k0 + kk < kmax && x0 + x4 + c < xmax) ? p[c] : 0.0f;
    } else {
      const int x = f >> 2, k4 = (f & 3) * 4;
      const float* p = M + (size_t)(x0 + x) * ld + k0 + k4;
      if (fast) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        reg[s * 4 + 0] = v.x; reg[s * 4 + 1] = v.y; reg[s * 4 + 2] = v.z; reg[s * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) reg[s * 4 + c] = (x0 + x < xmax && k0 + k4 + c < kmax) ? p[c] : 0.0f;
      }
    }
  }
}
template <bool MN_CONTIG>
__device__ __forceinline__ void store_tile(TileRow* S, const float (&reg)[8]) {
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int f = threadIdx.x + s * 256;
    if (MN_CONTIG) {
      const int kk = f >> 5, x4 = (f & 31) * 4;
      *reinterpret_cast<float4*>(&S[kk][x4]) = make_float4(reg[s * 4], reg[s * 4 + 1], reg[s * 4 + 2], reg[s * 4 + 3]);
    } else {
      const int x = f >> 2, k4 = (f & 3) * 4;
#pragma unroll
      for (int c = 0; c < 4; c++) S[k4 + c][x] = reg[s * 4 + c];
    }
  }
}

// Stages tile `kt` of one operand into ring slot S: asynchronously when it can, else into `reg` (stored by finish_tile).
template <bool MN_CONTIG>
__device__ __forceinline__ void begin_tile(const float* __restrict__ M, int ld, int x0, int k0, int xmax, int kmax, bool whole,
                                           TileRow* S, float (&reg)[8]) {
  if constexpr (MN_CONTIG) {
    if (whole) {
      tile_async<true>(M, ld, x0, k0, S);
      return;
    }
  }
  load_tile<MN_CONTIG>(M, ld, x0, k0, xmax, kmax, whole, reg);
}
template <bool MN_CONTIG>
__device__ __forceinline__ void finish_tile(bool whole, TileRow* S, const float (&reg)[8]) {
  if constexpr (MN_CONTIG) {
    if (whole) return;  // cp.async wrote it
  }
  store_tile<MN_CONTIG>(S, reg);
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 2) gemm_tiled_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ C, int m, int n, int k, int lda, int ldb,
                                                            int ldc, GemmEpilogue epi) {
  extern __shared__ __align__(16) float tiled_smem[];
  TileRow* As = reinterpret_cast<TileRow*>(tiled_smem);                          // As[stage * BK + kk][x]
  TileRow* Bs = reinterpret_cast<TileRow*>(tiled_smem) + (size_t)STAGES * BK;    // Bs[stage * BK + kk][x]

  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ri = threadIdx.x & 15, cj = threadIdx.x >> 4;

  // whole tile in range and 16-byte aligned: vector / asynchronous loads
  const bool a_al = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 4 == 0);
  const bool b_al = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 4 == 0);
  const bool a_ok = a_al && m0 + BM <= m, b_ok = b_al && n0 + BN <= n;

  float2 acc[8][4];  // acc[i][j2] = outputs (row i, columns 2 j2 and 2 j2 + 1) of the 8 x 8 micro-tile
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.0f, 0.0f);

  float ra[8], rb[8];
  const int ktiles = (k + BK - 1) / BK;
  // prologue: tiles 0 and 1 (one commit group per tile, empty groups keep the count uniform)
#pragma unroll
  for (int t = 0; t < STAGES - 1; t++) {
    if (t < ktiles) {
      const int k0 = t * BK;
      const bool kf = k0 + BK <= k;
      begin_tile<!TA>(A, lda, m0, k0, m, k, a_ok && kf, As + t * BK, ra);
      begin_tile<TB>(B, ldb, n0, k0, n, k, b_ok && kf, Bs + t * BK, rb);
      finish_tile<!TA>(a_ok && kf, As + t * BK, ra);
      finish_tile<TB>(b_ok && kf, Bs + t * BK, rb);
    }
    cp_async_commit();
  }
  cp_async_wait<STAGES - 2>();
  __syncthreads();

  for (int kt = 0; kt < ktiles; kt++) {
    const int cur = kt % STAGES, nxt = (kt + STAGES - 1) % STAGES;
    const int kn = kt + STAGES - 1;  // the tile staged during this iteration; its slot was last read in iteration kt - 1
    const int k0n = kn * BK;
    const bool more = kn < ktiles, kf = k0n + BK <= k;
    if (more) {
      begin_tile<!TA>(A, lda, m0, k0n, m, k, a_ok && kf, As + nxt * BK, ra);
      begin_tile<TB>(B, ldb, n0, k0n, n, k, b_ok && kf, Bs + nxt * BK, rb);
    }
    const TileRow* a_s = As + cur * BK;
    const TileRow* b_s = Bs + cur * BK;
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&a_s[kk][ri * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&a_s[kk][64 + ri * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&b_s[kk][cj * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&b_s[kk][64 + cj * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float2 aa = make_float2(a[i], a[i]);  // FFMA2 takes the scalar register as a broadcast operand: no move
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = __ffma2_rn(aa, b[j], acc[i][j]);
      }
    }
    if (more) {
      finish_tile<!TA>(a_ok && kf, As + nxt * BK, ra);
      finish_tile<TB>(b_ok && kf, Bs + nxt * BK, rb);
    }
    cp_async_commit();
    cp_async_wait<STAGES - 2>();  // everything but the newest group has landed: tile kt + 1 is complete
    __syncthreads();
  }

  const bool c_al = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (ldc % 4 == 0);
#pragma unroll
  for (int jh = 0; jh < 2; jh++)
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const int j = n0 + jh * 64 + cj * 4 + jj;
      if (j >= n) continue;
#pragma unroll
      for (int ih = 0; ih < 2; ih++) {
        const int i = m0 + ih * 64 + ri * 4;
        float* c = C + (size_t)j * ldc + i;
        float v[4];  // rows ih*4 .. ih*4+3 of micro-tile column jh*4 + jj
#pragma unroll
        for (int ii = 0; ii < 4; ii++) {
          const float2 pr = acc[ih * 4 + ii][(jh * 4 + jj) >> 1];
          v[ii] = ((jh * 4 + jj) & 1) ? pr.y : pr.x;
        }
        if (c_al && i + 4 <= m) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (epi.mode != 0) o = *reinterpret_cast<const float4*>(c);
          float4 w;
          w.x = epilogue_apply(epi, v[0], o.x);
          w.y = epilogue_apply(epi, v[1], o.y);
          w.z = epilogue_apply(epi, v[2], o.z);
          w.w = epilogue_apply(epi, v[3], o.w);
          *reinterpret_cast<float4*>(c) = w;
        } else {
#pragma unroll
          for (int ii = 0; ii < 4; ii++)
            if (i + ii < m) c[ii] = epilogue_apply(epi, v[ii], epi.mode != 0 ? c[ii] : 0.0f);
        }
      }
    }
}

template <bool TA, bool TB>
static void launch_t(const float* A, const float* B, float* C, int m, int n, int k, int lda, int ldb, int ldc, bool strict,
                     const GemmEpilogue& epi, cudaStream_t stream) {
  if (strict) {
    dim3 grid((m + 15) / 16, (n + 15) / 16);
    gemm_strict_kernel<TA, TB><<<grid, 256, 0, stream>>>(A, B, C, m, n, k, lda, ldb, ldc, epi);
  } else {
    static bool configured = false;  // one flag per <TA, TB> instantiation
    if (!configured) {
      CG_CUDA(cudaFuncSetAttribute(gemm_tiled_kernel<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILED_SMEM));
      configured = true;
    }
    dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN);
    gemm_tiled_kernel<TA, TB><<<grid, 256, TILED_SMEM, stream>>>(A, B, C, m, n, k, lda, ldb, ldc, epi);
  }
  CG_CUDA(cudaGetLastError());
}

void launch_gemm_f32(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb,
                     int ldc, bool strict, const GemmEpilogue& epi, cudaStream_t stream) {
  if (m == 0 || n == 0) return;
  if (ta) {
    if (tb) launch_t<true, true>(A, B, C, m, n, k, lda, ldb, ldc, strict, epi, stream);
    else launch_t<true, false>(A, B, C, m, n, k, lda, ldb, ldc, strict, epi, stream);
  } else {
    if (tb) launch_t<false, true>(A, B, C, m, n, k, lda, ldb, ldc, strict, epi, stream);
    else launch_t<false, false>(A, B, C, m, n, k, lda, ldb, ldc, strict, epi, stream);
  }
}

}  // namespace cg
