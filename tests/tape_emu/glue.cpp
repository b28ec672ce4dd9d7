// glue.cpp — TEST INFRASTRUCTURE (see shim.h). Compiled WITH the generated source (-include shim.h, then #include of
// the dumped file through -DTAPE_SOURCE=...), so that the generated `Ptrs` type and `tape_jit` are in scope.
#include <thread>
#include <vector>

#include TAPE_SOURCE

int run(int iters, float* trace, float** table) {
  Ptrs P;
  const size_t n = sizeof(P.p) / sizeof(P.p[0]);
  for (size_t i = 0; i < n; i++) P.p[i] = table[i];
  std::barrier<> bar(T);
  emu_barrier = &bar;
  std::vector<std::thread> threads;
  for (unsigned t = 0; t < T; t++)
    threads.emplace_back([=] {
      threadIdx.x = t;
      tape_jit(iters, trace, P);
    });
  for (auto& th : threads) th.join();
  return 0;
}
