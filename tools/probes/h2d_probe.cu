// H2D bandwidth probe: which host allocation gives the copy engine its full PCIe rate for 9 distinct 16 MiB inputs?
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t N = 16u << 20, K = 9;
  float* dev;
  CK(cudaMalloc(&dev, N * K));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (int mode = 0; mode < 5; mode++) {
    std::vector<char*> bufs(K);
    const char* name = "";
    char* slab = nullptr;
    for (size_t k = 0; k < K; k++) {
      if (mode == 0) { name = "cudaHostAlloc default, 9 buffers"; CK(cudaHostAlloc(&bufs[k], N, cudaHostAllocDefault)); }
      if (mode == 1) { name = "cudaHostAlloc write-combined, 9 buffers"; CK(cudaHostAlloc(&bufs[k], N, cudaHostAllocWriteCombined)); }
      if (mode == 2) { name = "malloc + cudaHostRegister, 9 buffers"; bufs[k] = (char*)aligned_alloc(4096, N); memset(bufs[k], 1, N); CK(cudaHostRegister(bufs[k], N, cudaHostRegisterDefault)); }
      if (mode == 3) {
        name = "2 MiB-aligned mmap + MADV_HUGEPAGE + cudaHostRegister, one slab";
        if (k == 0) {
          slab = (char*)mmap(nullptr, N * K + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
          slab = (char*)(((uintptr_t)slab + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
          madvise(slab, N * K, MADV_HUGEPAGE);
          memset(slab, 1, N * K);
          CK(cudaHostRegister(slab, N * K, cudaHostRegisterDefault));
        }
        bufs[k] = slab + k * N;
      }
      if (mode == 4) { name = "cudaHostAlloc default, ONE slab of 9 x 16 MiB"; if (k == 0) CK(cudaHostAlloc(&slab, N * K, cudaHostAllocDefault)); bufs[k] = slab + k * N; }
      if (mode != 2 && mode != 3) memset(bufs[k], 1, N);
    }
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaStreamSynchronize(st));
      double t0 = now();
      for (size_t k = 0; k < K; k++) CK(cudaMemcpyAsync((char*)dev + k * N, bufs[k], N, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      double dt = now() - t0;
      double t1 = now();
      for (size_t k = 0; k < K; k++) CK(cudaMemcpyAsync((char*)dev, bufs[0], N, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      double dt1 = now() - t1;
      printf("%-70s rep %d: 9 distinct %.2f ms (%.1f GB/s) | same buffer x9 %.2f ms (%.1f GB/s)\n", name, rep, dt * 1e3, N * K / dt / 1e9, dt1 * 1e3, N * K / dt1 / 1e9);
    }
    // D2H
    CK(cudaStreamSynchronize(st));
    double t0 = now();
    CK(cudaMemcpyAsync(bufs[0], dev, N, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    printf("%-70s D2H 16 MiB %.2f ms (%.1f GB/s)\n", name, (now() - t0) * 1e3, N / (now() - t0) / 1e9);
  }
  return 0;
}
