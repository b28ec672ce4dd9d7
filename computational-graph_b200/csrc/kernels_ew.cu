// kernels_ew.cu — fused elementwise program kernel, reductions, layout conversion.
#include "ew_interp.cuh"

namespace cg {

// ------------------------------------------------------------------------------------------------
// Fused elementwise program. HBM-bound by design: algorithmic bytes = 4 * n * (vector inputs + outputs).
// Tile = blockDim.x * VEC elements; thread t owns float4 #t and (VEC == 8) float4 #(t + blockDim.x) of
// the tile, so every warp-level access is one fully coalesced 512-byte request.
// ------------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 256;

template <int VEC, int MIN_BLOCKS>
__global__ void __launch_bounds__(EW_THREADS, MIN_BLOCKS) ew_program_kernel(const __grid_constant__ EwProgram p) {
  const size_t tile_elems = (size_t)EW_THREADS * VEC;
  const size_t stride = (size_t)EW_THREADS * 4;
  for (size_t tile = blockIdx.x; tile * tile_elems < p.n; tile += gridDim.x) {
    const size_t base = tile * tile_elems + (size_t)threadIdx.x * 4;
    if (base < p.n) ew_run<VEC>(p, base, stride);
  }
}

static int ew_vec_choice() {
  static int v = [] {
    const char* e = getenv("CG_EW_VEC");
    return e ? atoi(e) : 4;
  }();
  return v;
}

void launch_ew_program(const EwProgram& p, cudaStream_t stream) {
  if (p.n == 0) return;
  static bool configured = false;
  if (!configured) {
    // Same shared-memory carveout as the product kernel (which owns ~200 KB of the SM): a CTA of this kernel can
    // then be co-resident with a product CTA when the iteration graph runs them on parallel branches.
    int carve = cudaSharedmemCarveoutMaxShared;
    if (const char* e = getenv("CG_EW_CARVEOUT")) carve = atoi(e);
    if (carve >= -1) {
      cudaFuncSetAttribute(ew_program_kernel<4, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      cudaFuncSetAttribute(ew_program_kernel<8, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    configured = true;
  }
  // VEC = 4 (one float4 per operand per thread, <= 128 registers, two CTAs per SM — or one next to a product
  // CTA when the graph runs them concurrently); VEC = 8 halves the dispatch overhead per element but owns the SM.
  if (ew_vec_choice() == 8 && p.n > (size_t)EW_THREADS * 4 * kNumSMs) {
    size_t tiles = (p.n + EW_THREADS * 8 - 1) / (EW_THREADS * 8);
    size_t grid = tiles < (size_t)kNumSMs * 16 ? tiles : (size_t)kNumSMs * 16;
    ew_program_kernel<8, 1><<<(unsigned)grid, EW_THREADS, 0, stream>>>(p);
  } else {
    size_t tiles = (p.n + EW_THREADS * 4 - 1) / (EW_THREADS * 4);
    size_t grid = tiles < (size_t)kNumSMs * 16 ? tiles : (size_t)kNumSMs * 16;
    ew_program_kernel<4, 2><<<(unsigned)grid, EW_THREADS, 0, stream>>>(p);
  }
  CG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// reduce_sum: warp shuffle -> block -> ordered final pass by the last block (deterministic for a given
// n). Replaces the reference's one-atomicAdd-per-element kernel + cudaMalloc/D2H per call
// (src/gpu/ops.cu:143-153,190-205). workspace: [0, 1024) partials, [1024] ticket counter.
// ------------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1024;
size_t reduce_workspace_floats() { return RED_MAX_BLOCKS + 4; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0.0f;
  if (wid == 0) t = warp_sum(t);
  return t;  // valid in warp 0
}

__global__ void reduce_avg_strict_kernel(const float* __restrict__ src, size_t n, float denom, float* __restrict__ dst) {
  // cpu/ops.cpp:138-145 verbatim order: sequential f32 sum, then ops.rs:441 `/ size as f32`
  float sum = 0.0f;
  for (size_t i = 0; i < n; i++) sum = __fadd_rn(sum, src[i]);
  dst[0] = __fdiv_rn(sum, denom);
}

__global__ void __launch_bounds__(RED_THREADS) reduce_avg_kernel(const float* __restrict__ src, size_t n, float denom,
                                                                 float* __restrict__ dst, float* __restrict__ ws) {
  __shared__ float smem[32];
  __shared__ bool is_last;
  float acc = 0.0f;
  const size_t n4 = n / 4;
  const float4* src4 = reinterpret_cast<const float4*>(src);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = src4[i];
    acc += (v.x + v.y) + (v.z + v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) acc += src[n4 * 4 + threadIdx.x];
  float bs = block_sum(acc, smem);
  if (gridDim.x == 1) {
    if (threadIdx.x == 0) dst[0] = __fdiv_rn(bs, denom);
    return;
  }
  unsigned* ticket = reinterpret_cast<unsigned*>(ws + RED_MAX_BLOCKS);
  if (threadIdx.x == 0) {
    ws[blockIdx.x] = bs;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float a = 0.0f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += ws[i];
    __syncthreads();
    float tot = block_sum(a, smem);
    if (threadIdx.x == 0) {
      dst[0] = __fdiv_rn(tot, denom);
      *ticket = 0;  // re-arm for the next launch (CUDA-graph replay)
    }
  }
}

void launch_reduce_avg(const float* src, size_t n, float* dst, float* workspace, cudaStream_t stream) {
  launch_reduce_div(src, n, (float)n, dst, workspace, stream);
}

void launch_reduce_div(const float* src, size_t n, float denom, float* dst, float* workspace, cudaStream_t stream) {
  if (n <= kStrictReduceMax) {
    reduce_avg_strict_kernel<<<1, 1, 0, stream>>>(src, n, denom, dst);
    CG_CUDA(cudaGetLastError());
    return;
  }
  size_t blocks = (n / 4 + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > (size_t)kNumSMs * 4) blocks = (size_t)kNumSMs * 4;
  if (blocks > RED_MAX_BLOCKS) blocks = RED_MAX_BLOCKS;
  reduce_avg_kernel<<<(unsigned)blocks, RED_THREADS, 0, stream>>>(src, n, denom, dst, workspace);
  CG_CUDA(cudaGetLastError());
}

// all_equal / all_less_than (cpu/ops.cpp:118-136). `violation` must be zeroed by the caller's memset.
template <bool LESS>
__global__ void __launch_bounds__(256) all_cmp_kernel(const float* __restrict__ src, size_t n, float x,
                                                      int* __restrict__ violation) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = src[i];
    bad |= LESS ? (v >= x) : (v != x);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *violation = 1;
}

static void launch_all_cmp(bool less, const float* src, size_t n, float x, int* flag, cudaStream_t stream) {
  CG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
  if (n == 0) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  if (less) all_cmp_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(src, n, x, flag);
  else all_cmp_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(src, n, x, flag);
  CG_CUDA(cudaGetLastError());
}
void launch_all_equal(const float* src, size_t n, float x, int* flag, cudaStream_t s) { launch_all_cmp(false, src, n, x, flag, s); }
void launch_all_less_than(const float* src, size_t n, float x, int* flag, cudaStream_t s) { launch_all_cmp(true, src, n, x, flag, s); }

// Row-major host layout -> column-major storage (set_array(transpose = true), cpu/matrix.cpp:26-36).
__global__ void __launch_bounds__(256) rowmajor_to_colmajor_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                   unsigned h, unsigned w) {
  __shared__ float tile[32][33];
  const unsigned bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: column block, by: row block
  for (unsigned j = threadIdx.y; j < 32; j += blockDim.y) {
    unsigned row = by + j, col = bx + threadIdx.x;
    if (row < h && col < w) tile[j][threadIdx.x] = src[(size_t)row * w + col];
  }
  __syncthreads();
  for (unsigned j = threadIdx.y; j < 32; j += blockDim.y) {
    unsigned col = bx + j, row = by + threadIdx.x;
    if (row < h && col < w) dst[(size_t)col * h + row] = tile[threadIdx.x][j];
  }
}

void launch_rowmajor_to_colmajor(const float* src, float* dst, unsigned h, unsigned w, cudaStream_t stream) {
  if (h == 0 || w == 0) return;
  dim3 grid((w + 31) / 32, (h + 31) / 32), block(32, 8);
  rowmajor_to_colmajor_kernel<<<grid, block, 0, stream>>>(src, dst, h, w);
  CG_CUDA(cudaGetLastError());
}

}  // namespace cg
