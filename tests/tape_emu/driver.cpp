// driver.cpp — TEST INFRASTRUCTURE (see shim.h). Usage: emu <table file> <iters> <output file>
// Reads the pointer table the product dumped next to the generated source, gives every persistent buffer a host
// array with its initial contents, runs tape_jit on 512 threads, and writes the trace (iters x root_n) followed by
// the final contents of every table entry.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "shim.h"

thread_local EmuThreadIdx threadIdx;
std::barrier<>* emu_barrier = nullptr;
alignas(16) float S[64 * 1024];  // the CTA's dynamic shared memory (<= 224 KB)

struct Ptrs;  // defined by the generated source
extern "C" void tape_jit(int iters, float* trace, const Ptrs P);
extern const unsigned long long emu_ptrs_count;  // sizeof(Ptrs) / sizeof(float*), from glue.cpp

int run(int iters, float* trace, float** table);  // glue.cpp: builds a Ptrs from the table and starts the threads

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  auto get = [&]() {
    unsigned long long v = 0;
    if (fread(&v, sizeof v, 1, f) != 1) exit(4);
    return v;
  };
  const unsigned long long count = get(), root_off = get(), root_n = get(), root_is_leaf = get(), smem = get();
  (void)root_off, (void)root_is_leaf;
  if (smem > sizeof S / sizeof(float)) return 5;
  std::vector<std::vector<float>> bufs(count);
  std::vector<float*> table(count ? count : 1, nullptr);
  for (unsigned long long i = 0; i < count; i++) {
    get();  // address (the Python side matches leaves by it)
    const unsigned long long n = get();
    get(), get();  // offset, written
    bufs[i].resize(n + 4);  // float4 tails read (never use) up to 3 elements past n only when b + 4 <= n: no overrun; slack anyway
    if (fread(bufs[i].data(), sizeof(float), n, f) != n) return 6;
    table[i] = bufs[i].data();
  }
  fclose(f);
  const int iters = atoi(argv[2]);
  std::vector<float> trace((size_t)iters * root_n + 1, 0.0f);
  if (run(iters, trace.data(), table.data()) != 0) return 7;
  FILE* o = fopen(argv[3], "wb");
  if (!o) return 8;
  fwrite(trace.data(), sizeof(float), (size_t)iters * root_n, o);
  for (unsigned long long i = 0; i < count; i++) fwrite(bufs[i].data(), sizeof(float), bufs[i].size() - 4, o);
  fclose(o);
  return 0;
}
