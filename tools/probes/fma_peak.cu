// fma_peak.cu — measurement probe, NOT part of libcg_b200.so: the FP32 FMA-pipe peak of this GPU, for the roofline
// denominator of the FP32 SIMT product ("the build must measure both on the box").
// A register-resident loop of independent FFMA chains: 16 accumulators per thread, 1024 threads per CTA, two CTAs per SM.
// Built as tools/probes/libcg_probes.so by tools/probes/Makefile; bench.py loads it with ctypes when present.
#include <cuda_runtime.h>

#include <algorithm>

namespace {
constexpr int kChains = 16;
__global__ void __launch_bounds__(1024, 2) ffma_loop(float* out, int iters, float a, float b) {
  float acc[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) acc[c] = (float)(threadIdx.x + c);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) acc[c] = fmaf(acc[c], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < kChains; c++) s += acc[c];
  if (s == 12345.678f) out[0] = s;  // keeps the loop alive without a store in the common case
}
// The same loop with the packed FFMA2 of sm_100 (two IEEE fma.rn per instruction): half the issue slots per flop.
__global__ void __launch_bounds__(1024, 2) ffma2_loop(float* out, int iters, float a, float b) {
  float2 acc[kChains / 2];
#pragma unroll
  for (int c = 0; c < kChains / 2; c++) acc[c] = make_float2((float)(threadIdx.x + c), (float)c);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < kChains / 2; c++) acc[c] = __ffma2_rn(acc[c], aa, bb);
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < kChains / 2; c++) s += acc[c].x + acc[c].y;
  if (s == 12345.678f) out[0] = s;
}
}  // namespace

// TFLOP/s of the best of `reps` launches (each about `iters` x 16 FFMA per thread), on the current device.
static double probe(int iters, int reps, bool packed);
extern "C" double cg_probe_fma_tflops(int iters, int reps) { return probe(iters, reps, false); }
extern "C" double cg_probe_fma2_tflops(int iters, int reps) { return probe(iters, reps, true); }
static double probe(int iters, int reps, bool packed) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out = nullptr;
  if (cudaMalloc(&out, 4) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 2, threads = 1024;
  if (packed) ffma2_loop<<<grid, threads>>>(out, 64, 0.999f, 0.001f);  // warm-up
  else ffma_loop<<<grid, threads>>>(out, 64, 0.999f, 0.001f);
  double best = 0.0;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    if (packed) ffma2_loop<<<grid, threads>>>(out, iters, 0.999f, 0.001f);
    else ffma_loop<<<grid, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * kChains * (double)iters * threads * grid;
    best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}
