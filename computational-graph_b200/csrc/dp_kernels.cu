// dp_kernels.cu — the data-parallel exchange step as ONE kernel over NVLink peer memory.
//
// For every weight W the backward pass leaves a per-rank partial gradient g_r = X_r^T . E_r. Instead of
// ncclAllReduce(g) followed by an elementwise W <- W - lr * g pass, rank r
//   1. tells every peer that g_r is complete (the kernel starts only after the product that wrote it, stream order)
//      and waits until every peer has said the same                                   [flag barrier A]
//   2. for ITS 1/world slice of the matrix: loads the slice of g from every rank (peer loads over NVLink, ranks in
//      fixed order 0..world-1 so every replica is bit-identical), applies the reference's update
//      W <- W - (sum * lr) (two roundings, optimize.rs:149-152) once, and stores the new slice into the W of EVERY
//      rank (peer stores)                                                             [reduce-scatter + SGD + all-gather]
//   3. tells every peer that it has finished reading their g and writing their W, and the kernel does not complete
//      before every peer has said the same                                            [flag barrier B]
// Wire traffic per rank and weight: (world-1)/world of the matrix in each direction — what a ring all-reduce moves —
// but in two hops instead of 2 (world-1), with no separate update pass over HBM, and with no SM of the machine
// taken from the products except this kernel's own small grid.
//
// Flags are monotonically increasing epochs (one per launch, kept in device memory so CUDA-graph replays advance
// them); every rank launches these kernels in the same order on one in-order stream lane (exec.cpp), which is what
// makes "epoch >= mine" a correct wait. Waits are bounded: on timeout the kernel records an error and goes on, the
// host raises it after the next synchronisation — a lost rank must not hang the GPU.
#include "cg_common.h"

namespace cg {

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_stream(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// waits until flags[base + q] >= epoch for every rank q (thread q of the CTA polls rank q's flag). Once an exchange has
// timed out (err != 0) nothing waits any more: the remaining kernels of the iteration drain at once and the host
// raises the error after its next synchronisation (dp_check_errors).
__device__ __forceinline__ void wait_all(const uint32_t* own, int base, int world, uint32_t epoch, uint32_t* err,
                                         unsigned long long timeout_ns) {
  if ((int)threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(own + base + threadIdx.x) - epoch) < 0) {
      if (ld_acquire_sys(err) != 0) break;
      if (global_ns() - t0 > timeout_ns) {
        st_release_sys(err, 1u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// WORLD is a template parameter so that the loads of one trip — UNROLL float4 from each of the WORLD - 1 peers plus
// the local gradient and weight — are all issued before the first add: a peer load takes ~2000 cycles
// (B300_MICROARCH.md "NVLink": peer-LDG -> remote DRAM), so NVLink bandwidth is (bytes in flight) / latency: 775 GB/s x
// 1.06 us = 0.8 MB, and 64 CTAs x 512 threads x >= 2 peer loads x 16 B keeps >= 1 MB on the wire. The register
// budget is 64 per thread (__launch_bounds__(512, 2)): a CTA of this kernel must fit on an SM beside a CTA of the
// persistent product kernel (192 threads x 151 registers), or it would wait for a whole product to retire.
template <int WORLD, int UNROLL>
__global__ void __launch_bounds__(DP_THREADS, 2) dp_reduce_update_kernel(const DpReduceArgs a) {
  __shared__ uint32_t s_epoch;
  __shared__ bool s_last;
  uint32_t* own = a.flags[a.rank];
  if (threadIdx.x == 0) s_epoch = own[DP_FLAG_EPOCH] + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;

  // ---- barrier A: every rank's partial gradient is complete, and every rank is done reading the old weight ----
  if (blockIdx.x == 0 && (int)threadIdx.x < WORLD) st_release_sys(a.flags[threadIdx.x] + DP_FLAG_A + a.rank, epoch);
  wait_all(own, DP_FLAG_A, WORLD, epoch, own + DP_FLAG_ERR, a.timeout_ns);
  // A peer that never arrived: its gradient is not there to be summed. Leave W alone (the host raises the error at its
  // next synchronisation) but keep the ticket / epoch bookkeeping below going so that later kernels drain.
  __shared__ bool s_lost;
  if (threadIdx.x == 0) s_lost = ld_acquire_sys(own + DP_FLAG_ERR) != 0;
  __syncthreads();

  // ---- this rank's slice: sum over ranks in rank order, update once, store everywhere -------------------------
  const size_t n4 = a.n / 4;  // the caller guarantees n % 4 == 0 and 16-byte aligned buffers
  const size_t per = (n4 + WORLD - 1) / WORLD;
  const size_t lo = (size_t)a.rank * per, hi = s_lost ? 0 : (lo + per < n4 ? lo + per : n4);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * UNROLL) {
    float4 g[WORLD][UNROLL], w[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) ok[u] = i0 + (size_t)u * stride < hi;
#pragma unroll
    for (int r = 0; r < WORLD; r++)
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
        if (ok[u]) g[r][u] = ld_stream(a.g[r] + (i0 + (size_t)u * stride) * 4);
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
      if (ok[u]) w[u] = ld_stream(a.w[a.rank] + (i0 + (size_t)u * stride) * 4);
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (!ok[u]) continue;
      float4 sum = g[0][u];
#pragma unroll
      for (int r = 1; r < WORLD; r++) {
        sum.x = __fadd_rn(sum.x, g[r][u].x);
        sum.y = __fadd_rn(sum.y, g[r][u].y);
        sum.z = __fadd_rn(sum.z, g[r][u].z);
        sum.w = __fadd_rn(sum.w, g[r][u].w);
      }
      // optimize.rs:149-152: error *= lr; value -= error — two roundings, never an FMA
      float4 o;
      o.x = __fsub_rn(w[u].x, __fmul_rn(sum.x, a.lr));
      o.y = __fsub_rn(w[u].y, __fmul_rn(sum.y, a.lr));
      o.z = __fsub_rn(w[u].z, __fmul_rn(sum.z, a.lr));
      o.w = __fsub_rn(w[u].w, __fmul_rn(sum.w, a.lr));
#pragma unroll
      for (int r = 0; r < WORLD; r++) st_stream(a.w[r] + (i0 + (size_t)u * stride) * 4, o);
    }
  }

  // ---- barrier B: the last CTA of this rank signals, and waits for, the end of every rank's peer traffic -----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(own + DP_FLAG_TICKET, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < WORLD) st_release_sys(a.flags[threadIdx.x] + DP_FLAG_B + a.rank, epoch);
  wait_all(own, DP_FLAG_B, WORLD, epoch, own + DP_FLAG_ERR, a.timeout_ns);
  if (threadIdx.x == 0) {
    own[DP_FLAG_TICKET] = 0;
    own[DP_FLAG_EPOCH] = epoch;
    __threadfence();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same exchange with the sum done INSIDE the NVSwitch (NVLS): the gradient slot and the weight live in the symmetric
// multicast heap (symm.cpp), so for its 1/world slice a rank issues ONE multimem.ld_reduce per 16 bytes — the switch
// pulls the four floats from every GPU, adds them and returns the sum — applies W <- W - (sum * lr) (two roundings,
// optimize.rs:149-152) and issues ONE multimem.st that the switch replicates into every GPU's W. Per GPU and direction
// that is (1 + 1/world) n bytes on the wire instead of the 2 (world-1)/world n of the peer-memory kernel above: 1.125 n
// vs 1.75 n at 8 ranks (it is MORE at 2 ranks: 1.5 n vs n, which is why dp.cpp only selects it from 3 ranks up). Every
// element is summed once, by the switch, and broadcast, so the replicas stay bit-identical to each other; the order of
// the switch's f32 additions is the hardware's, which the 1e-5 / 1e-3 data-parallel tolerance of SURVEY §8e covers.
//
// Barriers: one multimem.red.add on a counter that exists on every GPU is "I have arrived" on all of them at once; a rank
// waits until its own copy reads epoch * world. The multicast address aliases the unicast one, hence fence.proxy.alias
// between the reduction and the polling loads.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mc_signal(uint32_t* mc_counter) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_counter), "r"(1u) : "memory");
  asm volatile("fence.proxy.alias;" ::: "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// thread 0 of the CTA polls this rank's counter until it reaches `target` (wrap-safe), bounded like wait_all
__device__ __forceinline__ void mc_wait(const uint32_t* counter, uint32_t target, uint32_t* err, unsigned long long timeout_ns) {
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(counter) - target) < 0) {
      if (ld_acquire_sys(err) != 0) break;
      if (global_ns() - t0 > timeout_ns) {
        st_release_sys(err, 1u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

template <int UNROLL>
__global__ void __launch_bounds__(DP_THREADS, 2) dp_multimem_update_kernel(const DpMcArgs a) {
  __shared__ uint32_t s_epoch;
  __shared__ bool s_last, s_lost;
  uint32_t* own = a.flags_uc;
  if (threadIdx.x == 0) s_epoch = own[DP_FLAG_EPOCH] + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const uint32_t target = epoch * (uint32_t)a.world;

  // ---- barrier A: every rank's partial gradient is complete, and every rank is done reading the old weight ----
  if (blockIdx.x == 0 && threadIdx.x == 0) mc_signal(a.flags_mc + DP_FLAG_A);
  mc_wait(own + DP_FLAG_A, target, own + DP_FLAG_ERR, a.timeout_ns);
  if (threadIdx.x == 0) s_lost = ld_acquire_sys(own + DP_FLAG_ERR) != 0;
  __syncthreads();

  // ---- this rank's slice: in-switch sum, one update, in-switch broadcast ---------------------------------------
  const size_t n4 = a.n / 4;
  const size_t per = (n4 + a.world - 1) / a.world;
  const size_t lo = (size_t)a.rank * per, hi = s_lost ? 0 : (lo + per < n4 ? lo + per : n4);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * UNROLL) {
    float4 g[UNROLL], w[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) ok[u] = i0 + (size_t)u * stride < hi;
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
      if (ok[u]) g[u] = mc_ld_reduce(a.g_mc + (i0 + (size_t)u * stride) * 4);
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
      if (ok[u]) w[u] = ld_stream(a.w_uc + (i0 + (size_t)u * stride) * 4);
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (!ok[u]) continue;
      float4 o;
      o.x = __fsub_rn(w[u].x, __fmul_rn(g[u].x, a.lr));
      o.y = __fsub_rn(w[u].y, __fmul_rn(g[u].y, a.lr));
      o.z = __fsub_rn(w[u].z, __fmul_rn(g[u].z, a.lr));
      o.w = __fsub_rn(w[u].w, __fmul_rn(g[u].w, a.lr));
      mc_st(a.w_mc + (i0 + (size_t)u * stride) * 4, o);
    }
  }

  // ---- barrier B: the last CTA of this rank signals, and waits for, the end of every rank's traffic ------------
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(own + DP_FLAG_TICKET, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) mc_signal(a.flags_mc + DP_FLAG_B);
  mc_wait(own + DP_FLAG_B, target, own + DP_FLAG_ERR, a.timeout_ns);
  if (threadIdx.x == 0) {
    own[DP_FLAG_TICKET] = 0;
    own[DP_FLAG_EPOCH] = epoch;
    __threadfence();
  }
}

template <int WORLD, int UNROLL>
void launch_world(const DpReduceArgs& a, cudaStream_t stream) {
  const size_t slice4 = (a.n / 4 + WORLD - 1) / WORLD;
  size_t grid = (slice4 + (size_t)DP_THREADS * UNROLL - 1) / ((size_t)DP_THREADS * UNROLL);
  if (grid > (size_t)dp_max_ctas()) grid = dp_max_ctas();
  if (grid < 1) grid = 1;
  dp_reduce_update_kernel<WORLD, UNROLL><<<(unsigned)grid, DP_THREADS, 0, stream>>>(a);
}

}  // namespace

int dp_max_ctas() {
  static const int v = [] {
    const char* e = getenv("CG_DP_CTAS");
    const int n = e ? atoi(e) : DP_DEFAULT_CTAS;
    return n < 1 ? 1 : n > kNumSMs ? kNumSMs : n;
  }();
  return v;
}

// The in-switch kernel only pulls n / world bytes and pushes n / world bytes per rank, so a small grid saturates the link
// (measured at 2 ranks: 32 CTAs and 64 CTAs take the same 190 us per 64 MB weight) and leaves the SMs to the products.
int dp_mc_max_ctas() {
  static const int v = [] {
    const char* e = getenv("CG_DP_MC_CTAS");
    const int n = e ? atoi(e) : DP_MC_DEFAULT_CTAS;
    return n < 1 ? 1 : n > kNumSMs ? kNumSMs : n;
  }();
  return v;
}

void launch_dp_multimem_update(const DpMcArgs& a, cudaStream_t stream) {
  if (a.n == 0) return;
  if (a.n % 4) throw Error("internal: fused data-parallel update needs a multiple of 4 elements");
  // The switch does the summing, so few threads keep the link busy: a rank only PULLS n / world bytes (the rest of its
  // inbound traffic is its peers' posted multicast stores). Two loads per thread from 3 ranks on, four at 2 ranks.
  const size_t slice4 = (a.n / 4 + a.world - 1) / a.world;
  static const int forced_unroll = [] {
    const char* e = getenv("CG_DP_MC_UNROLL");
    return e ? atoi(e) : 0;
  }();
  const int unroll = forced_unroll == 2 || forced_unroll == 4 ? forced_unroll : (a.world <= 2 ? 4 : 2);
  size_t grid = (slice4 + (size_t)DP_THREADS * unroll - 1) / ((size_t)DP_THREADS * unroll);
  if (grid > (size_t)dp_mc_max_ctas()) grid = dp_mc_max_ctas();
  if (grid < 1) grid = 1;
  if (unroll == 4) dp_multimem_update_kernel<4><<<(unsigned)grid, DP_THREADS, 0, stream>>>(a);
  else dp_multimem_update_kernel<2><<<(unsigned)grid, DP_THREADS, 0, stream>>>(a);
  CG_CUDA(cudaGetLastError());
}

void launch_dp_reduce_update(const DpReduceArgs& a, cudaStream_t stream) {
  if (a.n == 0) return;
  if (a.n % 4) throw Error("internal: fused data-parallel update needs a multiple of 4 elements");
  switch (a.world) {
    case 2: launch_world<2, 2>(a, stream); break;
    case 3: launch_world<3, 2>(a, stream); break;
    case 4: launch_world<4, 2>(a, stream); break;
    case 5: launch_world<5, 1>(a, stream); break;
    case 6: launch_world<6, 1>(a, stream); break;
    case 7: launch_world<7, 1>(a, stream); break;
    case 8: launch_world<8, 1>(a, stream); break;
    default: throw Error("internal: fused data-parallel update supports 2..8 ranks");
  }
  CG_CUDA(cudaGetLastError());
}

}  // namespace cg
