// gemm_tcgen05.cu — TF32 tensor-core product for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// C(m x n) = op(A) . op(B) on COLUMN-major f32 operands (the reference's storage, src/cpu/matrix.cpp:26-36),
// replacing the reference's cublasSgemm call (src/gpu/ops.cu:82-141). The three products of the hot path:
//   forward  X . W      (optimize.rs:135)  A = X   is M-contiguous ("MN-major"), B = W is K-contiguous ("K-major")
//   dX       E . W^T    (optimize.rs:197)  A = E   MN-major,                     B = W^T is N-contiguous (MN-major)
//   dW       X^T . E    (optimize.rs:198)  A = X^T K-major,                      B = E   K-major
// Both majors are fed to the tensor core as they lie in HBM — no transposition pass: the UMMA instruction
// descriptor carries a per-operand major bit, and the TMA tensor map is built so the tile lands in the
// canonical 128-byte-swizzled layout of that major:
//   K-major  : 2-D map {K, MN}, box {32, ROWS}            -> smem [ROWS][32 floats], 8-row swizzle atoms
//   MN-major : 3-D map {32, K, MN/32}, box {32, 32, ROWS/32} -> smem [ROWS/32][32 k-rows][32 floats], swizzled in
//              32-byte granules (SWIZZLE_128B_ATOM_32B / UMMA SWIZZLE_128B_BASE32B: the only layout from which
//              the tensor core transposes 32-bit operands)
// One persistent CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 =
// epilogue (TMEM -> registers -> coalesced column stores, with the SGD update or the error accumulation fused).
// 4-stage smem ring (full/empty mbarriers), 2 accumulator stages in TMEM so the epilogue of tile i overlaps
// the main loop of tile i+1. UMMA shape 128 x BN x 8 (kind::tf32), BN = 256 or 128.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <unordered_map>

#include "cg_common.h"

namespace cg {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;          // 32 floats = one 128-byte swizzle row
constexpr int UMMA_K = 8;       // kind::tf32: 32 bytes of K per instruction
constexpr int ACC_STAGES = 2;
constexpr int NUM_THREADS = 192;

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"): start address, leading /
// stride byte offsets (all >> 4), descriptor version 1 (Blackwell), swizzle layout type.
//   K-major  f32 tile: SWIZZLE_128B, atoms of 8 rows x 128 B; SBO = 1024 B between 8-row groups.
//   MN-major f32 tile: 32-bit operands can only be transposed by the tensor core from the
//     SWIZZLE_128B_BASE32B layout (32-byte swizzle granules; atoms of 4 k-rows x 128 B): LBO = distance between
//     32-element MN chunks, SBO = distance between 4-row k groups. TMA produces it with SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

// Persistent grid size. Data-parallel runs leave a few SMs to the collective kernels (dp.cpp), so that an all-reduce
// does not wait for a whole persistent product to drain before it can start.
int g_max_ctas = kNumSMs;

// Shared-memory descriptor strides of the two canonical operand layouts (bytes).
constexpr uint32_t K_LBO = 16, K_SBO = 1024, K_STEP = UMMA_K * 4;   // K-major: 8-row atoms 1024 B apart; +32 B per K step
constexpr uint32_t MN_LBO = BK * 128, MN_SBO = 512, MN_STEP = 1024;  // MN-major: 32-wide chunks 4096 B apart; 4-row atoms; +8 k-rows per step
constexpr uint32_t MN_LAYOUT = 1;                                    // SWIZZLE_128B_BASE32B

template <int BN, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     float* __restrict__ C, int m, int n, int k, int ldc, int epi_mode, float lr) {
  constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;  // 512 or 256: a power of two >= 32
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = m / BM, tiles_n = n / BN, num_tiles = tiles_m * tiles_n, num_kb = k / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < ACC_STAGES; s++) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (A_MN) tma_load_3d(sa, &map_a, &full_bar[stage], 0, kb * BK, m0 / 32);
          else tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m0);
          if (B_MN) tma_load_3d(sb, &map_b, &full_bar[stage], 0, kb * BK, n0 / 32);
          else tma_load_2d(sb, &map_b, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, per-operand major bits, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; kk++) {
            const uint64_t da = A_MN ? make_smem_desc(sa + kk * MN_STEP, MN_LBO, MN_SBO, MN_LAYOUT)
                                     : make_smem_desc(sa + kk * K_STEP, K_LBO, K_SBO, 2);
            const uint64_t db = B_MN ? make_smem_desc(sb + kk * MN_STEP, MN_LBO, MN_SBO, MN_LAYOUT)
                                     : make_smem_desc(sb + kk * K_STEP, K_LBO, K_SBO, 2);
            tcgen05_mma_tf32(tmem_d, da, db, idesc, (kb | kk) != 0);
          }
          tcgen05_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        tcgen05_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (column-major C: lanes = consecutive rows => coalesced) =====
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      float* crow = C + (size_t)n0 * ldc + m0 + quad * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; c++) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + c * 32 + ((uint32_t)(quad * 32) << 16), v);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          // every value of this accumulator stage is in registers: hand TMEM back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        float* cp = crow + (size_t)(c * 32) * ldc;
        if (epi_mode == 0) {
#pragma unroll
          for (int j = 0; j < 32; j++) cp[(size_t)j * ldc] = __uint_as_float(v[j]);
        } else if (epi_mode == 1) {
          // fused SGD (optimize.rs:149-152): W <- W - (g * lr), two roundings
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; j++) old[j] = cp[(size_t)j * ldc];
#pragma unroll
          for (int j = 0; j < 32; j++) cp[(size_t)j * ldc] = __fsub_rn(old[j], __fmul_rn(__uint_as_float(v[j]), lr));
        } else {
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; j++) old[j] = cp[(size_t)j * ldc];
#pragma unroll
          for (int j = 0; j < 32; j++) cp[(size_t)j * ldc] = __fadd_rn(old[j], __uint_as_float(v[j]));
        }
      }
      if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// =====================================================================================================================
// CTA-pair kernel (cta_group::2): two SMs of one TPC compute a 256 x W tile together. CTA r of the pair stages rows
// [m0 + 128 r, +128) of A and columns [n0 + r W/2, +W/2) of B, so each SM fills 16 KB + <= 16 KB of shared memory per
// 32-deep k-block instead of 16 KB + 32 KB — the single-CTA kernel above is bound by that fill (its BN = 128 variant,
// which needs 4/3 of the bytes per flop, runs 20-37 % slower: tools/gemm_bench.py with CG_TF32_BN128=1). The leader
// CTA's elected thread issues every tcgen05.mma for the pair; both accumulators (128 lanes x W columns each) stay in
// their own SM's TMEM and each CTA's epilogue warps write their 128 rows of C.
//
// Work is balanced by COLUMN RANGES, not whole tiles: the (m / 256) row-blocks x n columns are laid end to end and
// every pair takes one contiguous range of U columns, cut into chunks of <= 256 columns that do not cross a row-block.
// 1024 x 4096: 16384 columns over 74 pairs -> U = 224 (73.1 pairs busy) instead of 64 tiles of 256 on 74 pairs.
//
// Barriers (same offsets in both CTAs):  full[s]   leader only, 2 arrivals (one producer per CTA) + the bytes of both
//                                        empty[s]  per CTA, released by the leader's multicast tcgen05.commit
//                                        tfull[a]  per CTA, multicast commit after the last k-block of a chunk
//                                        tempty[a] leader only, 8 arrivals (4 epilogue warps x 2 CTAs)
// =====================================================================================================================
constexpr int STAGES2 = 6;
constexpr uint32_t A_BYTES2 = BM * BK * 4;          // 128 rows of A per CTA
constexpr uint32_t B_BYTES2 = 128 * BK * 4;         // <= 128 columns of B per CTA
constexpr uint32_t STAGE_BYTES2 = A_BYTES2 + B_BYTES2;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Arrival on a barrier of the peer CTA that publishes SHARED-memory writes to the tensor core (the splitters' lo tiles):
// the writes were already made visible to the async proxy by fence.proxy.async, so the arrival needs no cluster-scope
// release — `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR and cost the splitters half their time (ncu source
// page, profiles/r02l_*). Same form as CUTLASS's ClusterBarrier::arrive(cta_id).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                      uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// What the tensor core drops when it truncates an f32 operand to TF32: exact, two instructions per element (the splitters
// are issue-bound). An infinite operand gives lo = inf - inf = nan, so a product that plain f32 arithmetic would have
// made inf comes out nan — non-finite either way.
__device__ __forceinline__ float tf32_lo(float x) {
  return __fsub_rn(x, __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}

// The tile sequence of one pair (static round-robin over (row-block, column-tile)): every role of both CTAs walks
// the same list. Consecutive pairs take consecutive column tiles of one row-block, so they share its A rows in L2.
// (Tried and dropped, profiles/r02m_tile_order_prefetch.txt: taking the row-blocks in bands so that B streams past an
// L2-resident A panel, and L2 prefetch of later k-blocks — the rate does not move (prefetch costs 3-6 %): the kernels
// run against the power cap, not against operand latency.)
struct ChunkWalk {
  int t, ntiles, stride, tiles_n, wmax;
  __device__ __forceinline__ ChunkWalk(int pair, int npairs, int total_cols, int n, int wmax_) : wmax(wmax_) {
    tiles_n = n / wmax_;
    ntiles = total_cols / wmax_;
    t = pair;
    stride = npairs;
  }
  // rb: 256-row block, n0: first column, w: width (<= 256)
  __device__ __forceinline__ bool next(int& rb, int& n0, int& w) {
    if (t >= ntiles) return false;
    rb = t / tiles_n;
    n0 = (t - rb * tiles_n) * wmax;
    w = wmax;
    t += stride;
    return true;
  }
};

// SPLIT3 — FP32-accurate products on the tensor core ("3xTF32"). The tensor core truncates every f32 operand to its top 10
// mantissa bits (hi); the bits it drops, lo = x - hi, are themselves an f32 (exact). Six more warps per CTA (the
// "splitters") compute lo for every staged tile, in shared memory, right behind the TMA: they wait on the CTA's own full
// barrier, read the stage, write lo at the same (swizzled) position of a second tile behind it, make the writes visible to
// the async proxy and arrive on the LEADER's split barrier, which is what the MMA thread waits on. Per k-step three MMAs:
//   A_lo.B + A.B_lo  ->  the LO accumulator          A.B  ->  the HI accumulator
//   = lo_a' hi_b + hi_a lo_b'                             = hi_a hi_b           (primes: lo truncated to ITS top 10 bits)
// which leaves out lo_a lo_b (2^-22 of a term on average) and the second truncations (< 2^-21): a product term is good to
// ~1e-6. Nothing is materialised in HBM and the operands cross L2 -> shared memory once; a stage holds four tiles.
//
// The tensor core's own accumulation is the other half of the problem: every tcgen05.mma adds its 8-deep partial sum to
// the f32 accumulator with TRUNCATION, half an ulp of the accumulator lost towards zero per instruction — measured on this
// part: k = 4096 through one accumulator ends 770 ulps short (1.5e-3 on a value of 30). So (a) the small terms have an
// accumulator of their own, 2^-11 of the size, whose truncations do not matter, and (b) the contraction is cut into
// chunks of SPLIT_KC k-blocks (256 deep: 32 truncating adds), each into a fresh accumulator stage, and the chunks are
// summed by the epilogue warps in registers with round-to-nearest adds while the next chunk is on the tensor core. That
// needs hi + lo for two stages in the 512 TMEM columns, hence column tiles of <= 128 here, and the running sum in
// registers: eight epilogue warps, 64 sums per thread. Error of a k = 4096 product: ~16 ulps, what a k-ascending f32 FMA loop gives.
//
// Barriers in SPLIT3:                     full[s]   per CTA, its own producer's arrival + its own bytes
//                                        split[s]  leader only, 12 arrivals (6 splitter warps x 2 CTAs)
//                                        tfull[a] / tempty[a]  once per CHUNK instead of once per tile
constexpr int SPLIT_KC = 8;
constexpr int SPLIT_STAGES = 4;
constexpr int SPLIT_EPI_WARPS = 8;    // warps 2-9: two per TMEM lane quad, 64 columns of the tile each (64 running sums per thread)
constexpr int SPLIT_SPLIT_WARPS = 6;  // warps 10-15
constexpr uint32_t SPLIT_TILE_BYTES = A_BYTES2 + 64 * BK * 4;  // 128 rows of A + this CTA's <= 64 columns of B: 24 KB
constexpr int PAIR_THREADS = 192, PAIR_THREADS_SPLIT3 = 512;

template <bool A_MN, bool B_MN, bool SPLIT3>
__global__ void __launch_bounds__(SPLIT3 ? PAIR_THREADS_SPLIT3 : PAIR_THREADS, 1)
    gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_big,
                          const __grid_constant__ CUtensorMap map_b_small, float* __restrict__ C, int m, int n, int k, int ldc,
                          int epi_mode, float lr, int wmax) {
  constexpr int NST = SPLIT3 ? SPLIT_STAGES : STAGES2;
  constexpr uint32_t HB = SPLIT3 ? SPLIT_TILE_BYTES : STAGE_BYTES2;  // [A][B]: what the TMA stages
  constexpr uint32_t SB = SPLIT3 ? 2 * HB : HB;                      // SPLIT3: [A][B][A_lo][B_lo]
  // B arrives in one box of wmax / 2 columns for a full-width chunk, in 16-row (K-major) / 32-column (MN-major) boxes for
  // the partial chunks at the ends of a range: TMA issue cost is per instruction, so the common case must be one box.
  constexpr int SMALL = B_MN ? 32 : 16;
  const int BIG = wmax >> 1;
  constexpr uint32_t TMEM_COLS = 512;  // 2 accumulator stages of <= 256 columns
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NST * SB);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* tfull_bar = empty_bar + NST;
  uint64_t* tempty_bar = tfull_bar + ACC_STAGES;
  uint64_t* split_bar = tempty_bar + ACC_STAGES;  // SPLIT3 only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(split_bar + NST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int total_cols = (m / 256) * n, num_kb = k / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_big) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_small) : "memory");
    for (int s = 0; s < NST; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 2 * SPLIT_SPLIT_WARPS);
    }
    for (int s = 0; s < ACC_STAGES; s++) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], SPLIT3 ? 2 * SPLIT_EPI_WARPS : 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // one warp of EACH CTA of the pair performs the 2-CTA allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();  // barrier inits and both allocations are visible to the peer before anything is signalled
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one per CTA): its own half of the operands, signalling the LEADER's full barrier =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      ChunkWalk walk(pair, (int)(gridDim.x >> 1), total_cols, n, wmax);
      int rb, n0, w;
      while (walk.next(rb, n0, w)) {
        const int m0 = rb * 256 + (int)rank * BM;
        const int half = w >> 1, b0 = n0 + (int)rank * half;
        const uint32_t bytes = A_BYTES2 + (uint32_t)half * BK * 4;
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SB;
          uint8_t* sb = sa + A_BYTES2;
          // SPLIT3: the bytes are counted where they land (the CTA's splitters wait for them); otherwise on the leader
          const uint32_t lbar = map_to_cta(smem_u32(&full_bar[stage]), SPLIT3 ? rank : 0);
          if (A_MN) tma2_load_3d(sa, &map_a, lbar, 0, kb * BK, m0 / 32);
          else tma2_load_2d(sa, &map_a, lbar, kb * BK, m0);
          int done = 0;
          if (B_MN) {
            for (; done + BIG <= half; done += BIG) tma2_load_3d(sb + done * 128, &map_b_big, lbar, 0, kb * BK, (b0 + done) / 32);
            for (; done < half; done += SMALL) tma2_load_3d(sb + done * 128, &map_b_small, lbar, 0, kb * BK, (b0 + done) / 32);
          } else {
            for (; done + BIG <= half; done += BIG) tma2_load_2d(sb + done * 128, &map_b_big, lbar, kb * BK, b0 + done);
            for (; done < half; done += SMALL) tma2_load_2d(sb + done * 128, &map_b_small, lbar, kb * BK, b0 + done);
          }
          // the leader accounts for the bytes of both CTAs (its single arrival + 2 x bytes complete the phase)
          if (SPLIT3) mbar_expect_tx(&full_bar[stage], bytes);
          else if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * bytes);
          if (++stage == NST) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both tensor cores =====
    if (rank == 0 && lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      ChunkWalk walk(pair, (int)(gridDim.x >> 1), total_cols, n, wmax);
      int rb, n0, w;
      while (walk.next(rb, n0, w)) {
        // D = F32, A = B = TF32, major bits, N >> 3, M = 256 >> 4
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(w >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        if (!SPLIT3) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // both epilogues have drained this accumulator stage
          tcgen05_fence_after();
        }
        for (int kb = 0; kb < num_kb; kb++) {
          const bool chunk_start = SPLIT3 && kb % SPLIT_KC == 0;
          if (chunk_start) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tcgen05_fence_after();
          }
          const uint32_t tmem_d = tmem_base + acc * 256;  // SPLIT3: hi accumulator; lo accumulator 128 columns behind
          mbar_wait(SPLIT3 ? &split_bar[stage] : &full_bar[stage], phase);  // both CTAs' bytes have landed (and are split)
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SB), sb = sa + A_BYTES2;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; kk++) {
            const uint64_t da = A_MN ? make_smem_desc(sa + kk * MN_STEP, MN_LBO, MN_SBO, MN_LAYOUT)
                                     : make_smem_desc(sa + kk * K_STEP, K_LBO, K_SBO, 2);
            const uint64_t db = B_MN ? make_smem_desc(sb + kk * MN_STEP, MN_LBO, MN_SBO, MN_LAYOUT)
                                     : make_smem_desc(sb + kk * K_STEP, K_LBO, K_SBO, 2);
            if (SPLIT3) {
              const uint32_t la = sa + HB, lb = sb + HB;
              const uint64_t da_lo = A_MN ? make_smem_desc(la + kk * MN_STEP, MN_LBO, MN_SBO, MN_LAYOUT)
                                          : make_smem_desc(la + kk * K_STEP, K_LBO, K_SBO, 2);
              const uint64_t db_lo = B_MN ? make_smem_desc(lb + kk * MN_STEP, MN_LBO, MN_SBO, MN_LAYOUT)
                                          : make_smem_desc(lb + kk * K_STEP, K_LBO, K_SBO, 2);
              const uint32_t fresh = (chunk_start && kk == 0) ? 0u : 1u;
              tcgen05_mma_tf32_pair(tmem_d + 128, da_lo, db, idesc, fresh);
              tcgen05_mma_tf32_pair(tmem_d + 128, da, db_lo, idesc, 1);
              tcgen05_mma_tf32_pair(tmem_d, da, db, idesc, fresh);
            } else {
              tcgen05_mma_tf32_pair(tmem_d, da, db, idesc, (kb | kk) != 0);
            }
          }
          tcgen05_commit_pair(&empty_bar[stage]);  // frees the slot in BOTH CTAs once these MMAs have read it
          if (++stage == NST) stage = 0, phase ^= 1;
          if (SPLIT3 && (kb % SPLIT_KC == SPLIT_KC - 1 || kb == num_kb - 1)) {
            tcgen05_commit_pair(&tfull_bar[acc]);  // this chunk's accumulators complete -> both epilogues
            if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
          }
        }
        if (!SPLIT3) {
          tcgen05_commit_pair(&tfull_bar[acc]);  // accumulators complete -> both epilogues
          if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
        }
      }
    }
  } else if (SPLIT3 && warp >= 2 + SPLIT_EPI_WARPS) {
    // ===== splitters: lo = x - (x truncated to TF32) for every element of the stage, written behind it =====
    constexpr int NT = SPLIT_SPLIT_WARPS * 32;
    constexpr int PER = (int)(SPLIT_TILE_BYTES / 16 + NT - 1) / NT;  // 16-byte vectors per thread and stage: 8
    const int t = (int)threadIdx.x - (2 + SPLIT_EPI_WARPS) * 32;
    uint32_t stage = 0, phase = 0;
    ChunkWalk walk(pair, (int)(gridDim.x >> 1), total_cols, n, wmax);
    int rb, n0, w;
    while (walk.next(rb, n0, w)) {
      const int vecs = (int)(A_BYTES2 + (uint32_t)(w >> 1) * BK * 4) >> 4;  // A tile and this CTA's half of B are contiguous
      for (int kb = 0; kb < num_kb; kb++) {
        mbar_wait(&full_bar[stage], phase);
        const float4* src = reinterpret_cast<const float4*>(smem + stage * SB);
        float4* dst = reinterpret_cast<float4*>(smem + stage * SB + HB);
        float4 x[PER];
#pragma unroll
        for (int u = 0; u < PER; u++)
          if (t + u * NT < vecs) x[u] = src[t + u * NT];
#pragma unroll
        for (int u = 0; u < PER; u++)
          if (t + u * NT < vecs) {
            float4 lo;
            lo.x = tf32_lo(x[u].x), lo.y = tf32_lo(x[u].y), lo.z = tf32_lo(x[u].z), lo.w = tf32_lo(x[u].w);
            dst[t + u * NT] = lo;
          }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&split_bar[stage]), 0));
        if (++stage == NST) stage = 0, phase ^= 1;
      }
    }
  } else if (SPLIT3) {
    // ===== epilogue, SPLIT3: the chunks of the contraction summed in registers (round to nearest), then the stores.
    // Eight warps: warp e works on TMEM lanes [32 (warp % 4), +32) — the only ones it may touch — and columns [64 (e / 4), +64)
    const int quad = warp & 3, col0 = ((warp - 2) >> 2) * 64;
    uint32_t acc = 0, acc_phase = 0;
    ChunkWalk walk(pair, (int)(gridDim.x >> 1), total_cols, n, wmax);
    int rb, n0, w;
    const int nchunk = (num_kb + SPLIT_KC - 1) / SPLIT_KC;
    while (walk.next(rb, n0, w)) {
      const int m0 = rb * 256 + (int)rank * BM;
      float sum[64];
      for (int ch = 0; ch < nchunk; ch++) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t t0 = tmem_base + acc * 256 + col0 + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          if (col0 + c * 16 < w) {
            uint32_t hi[16], lo[16];
            tmem_ld_32x32b_x16(t0 + c * 16, hi);
            tmem_ld_32x32b_x16(t0 + 128 + c * 16, lo);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const float part = __fadd_rn(__uint_as_float(hi[j]), __uint_as_float(lo[j]));
              sum[c * 16 + j] = ch == 0 ? part : __fadd_rn(sum[c * 16 + j], part);
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&tempty_bar[acc]);
          else mbar_arrive_remote(map_to_cta(smem_u32(&tempty_bar[acc]), 0));  // once per chunk: no cluster-scope membar
        }
        if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
      }
      float* crow = C + (size_t)(n0 + col0) * ldc + m0 + quad * 32 + lane;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        if (col0 + c * 16 < w) {
          float* cp = crow + (size_t)(c * 16) * ldc;
          if (epi_mode == 0) {
#pragma unroll
            for (int j = 0; j < 16; j++) cp[(size_t)j * ldc] = sum[c * 16 + j];
          } else {
            float old[16];
#pragma unroll
            for (int j = 0; j < 16; j++) old[j] = cp[(size_t)j * ldc];
#pragma unroll
            for (int j = 0; j < 16; j++)
              cp[(size_t)j * ldc] = epi_mode == 1 ? __fsub_rn(old[j], __fmul_rn(sum[c * 16 + j], lr)) : __fadd_rn(old[j], sum[c * 16 + j]);
          }
        }
      }
    }
  } else {
    // ===== epilogue (each CTA: its own 128 rows): TMEM -> registers -> coalesced column stores =====
    const int quad = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    ChunkWalk walk(pair, (int)(gridDim.x >> 1), total_cols, n, wmax);
    int rb, n0, w;
    while (walk.next(rb, n0, w)) {
      const int m0 = rb * 256 + (int)rank * BM;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      float* crow = C + (size_t)n0 * ldc + m0 + quad * 32 + lane;
      const int nchunks = w >> 5;
#pragma unroll 1
      for (int c = 0; c < nchunks; c++) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + acc * 256 + c * 32 + ((uint32_t)(quad * 32) << 16), v);
        tmem_ld_wait();
        if (c == nchunks - 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&tempty_bar[acc]);
            else mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[acc]), 0));
          }
        }
        float* cp = crow + (size_t)(c * 32) * ldc;
        if (epi_mode == 0) {
#pragma unroll
          for (int j = 0; j < 32; j++) cp[(size_t)j * ldc] = __uint_as_float(v[j]);
        } else if (epi_mode == 1) {
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; j++) old[j] = cp[(size_t)j * ldc];
#pragma unroll
          for (int j = 0; j < 32; j++) cp[(size_t)j * ldc] = __fsub_rn(old[j], __fmul_rn(__uint_as_float(v[j]), lr));
        } else {
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; j++) old[j] = cp[(size_t)j * ldc];
#pragma unroll
          for (int j = 0; j < 32; j++) cp[(size_t)j * ldc] = __fadd_rn(old[j], __uint_as_float(v[j]));
        }
      }
      if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
    }
  }

  // neither CTA may leave (or free TMEM) while the peer can still signal its barriers or read its shared memory
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

void resolve_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so that libcg_b200.so carries no link-time dependency on libcuda
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  if (!g_encode) throw Error("cuTensorMapEncodeTiled is not available from this driver");
}

constexpr CUtensorMapSwizzle MN_TMA_SWIZZLE = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;

// Tensor map of one operand tile. `mn` = extent along M (A) or N (B); `kdim` = contraction extent;
// `ld` = stored leading dimension (floats); rows = tile extent along mn.
CUtensorMap make_operand_map(const float* base, bool mn_major, int mn, int kdim, int ld, int rows) {
  resolve_encode();
  CUtensorMap map;
  CUresult r;
  if (!mn_major) {
    // contiguous along K: element (x, kk) at base[x * ld + kk]
    cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)mn};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    r = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // contiguous along MN: element (x, kk) at base[kk * ld + x]; viewed as {32, K, MN/32} so that one box
    // lands as [MN/32 chunks][BK k-rows][32 floats] — the canonical MN-major SWIZZLE_128B layout
    cuuint64_t dims[3] = {32, (cuuint64_t)kdim, (cuuint64_t)(mn / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)(rows / 32)};
    cuuint32_t estr[3] = {1, 1, 1};
    r = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, MN_TMA_SWIZZLE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return map;
}

template <int BN, bool A_MN, bool B_MN>
void launch_variant(const CUtensorMap& ma, const CUtensorMap& mb, float* C, int m, int n, int k, int ldc,
                    const GemmEpilogue& epi, cudaStream_t stream) {
  constexpr int STAGES = BN == 256 ? 4 : 6;  // 192 KB of operands in flight either way
  constexpr size_t smem = (size_t)STAGES * (BM + BN) * BK * 4 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    CG_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int tiles = (m / BM) * (n / BN);
  const int grid = tiles < g_max_ctas ? tiles : g_max_ctas;
  gemm_tf32_kernel<BN, A_MN, B_MN, STAGES><<<grid, NUM_THREADS, smem, stream>>>(ma, mb, C, m, n, k, ldc, epi.mode, epi.lr);
  CG_CUDA(cudaGetLastError());
}

template <int BN>
void launch_bn(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb, int ldc,
               const GemmEpilogue& epi, cudaStream_t stream) {
  // op(A)(i, kk): not transposed => A[kk * lda + i] (M-contiguous, MN-major); transposed => A[i * lda + kk] (K-major)
  // op(B)(kk, j): not transposed => B[j * ldb + kk] (K-major);                transposed => B[kk * ldb + j] (MN-major)
  const bool a_mn = !ta, b_mn = tb;
  const CUtensorMap ma = make_operand_map(A, a_mn, m, k, lda, BM);
  const CUtensorMap mb = make_operand_map(B, b_mn, n, k, ldb, BN);
  if (a_mn && !b_mn) launch_variant<BN, true, false>(ma, mb, C, m, n, k, ldc, epi, stream);
  else if (a_mn && b_mn) launch_variant<BN, true, true>(ma, mb, C, m, n, k, ldc, epi, stream);
  else if (!a_mn && !b_mn) launch_variant<BN, false, false>(ma, mb, C, m, n, k, ldc, epi, stream);
  else launch_variant<BN, false, true>(ma, mb, C, m, n, k, ldc, epi, stream);
}

template <bool A_MN, bool B_MN, bool SPLIT3>
void launch_pair_variant(const CUtensorMap& ma, const CUtensorMap& mb_big, const CUtensorMap& mb_small, float* C, int m, int n,
                         int k, int ldc, const GemmEpilogue& epi, int wmax, int pairs, cudaStream_t stream) {
  constexpr size_t smem = (size_t)(SPLIT3 ? SPLIT_STAGES * 2 * SPLIT_TILE_BYTES : STAGES2 * STAGE_BYTES2) + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    CG_CUDA(cudaFuncSetAttribute(gemm_tf32_pair_kernel<A_MN, B_MN, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(SPLIT3 ? PAIR_THREADS_SPLIT3 : PAIR_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CG_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_pair_kernel<A_MN, B_MN, SPLIT3>, ma, mb_big, mb_small, C, m, n, k, ldc, epi.mode, epi.lr,
                             wmax));
}

// On by default (cg_set_tf32_pair(0) selects the single-CTA kernel for every shape). A dense GEMM on this 1 kW part runs
// against the power cap, so what decides the sustained rate is energy per flop: the pair kernel fills 32 KB of shared
// memory per SM and k-block instead of 48 KB and reads 64 instead of 96 B/clk of operands, and with both kernels held
// at ~990 W for 1.5 s it sustains 639 / 640 / 613 TFLOP/s on the three 8192-row products against 560 / 563 / 581 for
// the single-CTA kernel (cuBLAS TF32: 664 / 661 / 639), 574 / 574 / 541 against 542 / 554 / 510 at 1024 rows (cuBLAS:
// 580 / 576 / 560) — tools/gemm_vs_cublas.py, profiles/r02c_gemm_vs_cublas_pair{0,1}.jsonl. Round 1 compared best-of-N
// bursts, where the clocks have not yet come down and the two kernels tie.
int g_pair_enabled = 1;
bool pair_kernel_enabled() { return g_pair_enabled != 0; }

bool pair_shape(int m, int n, int k) { return m % 256 == 0 && n % 64 == 0 && k % BK == 0; }
bool pair_eligible(int m, int n, int k) { return pair_kernel_enabled() && pair_shape(m, n, k); }

template <bool SPLIT3>
void launch_pair(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb, int ldc,
                 const GemmEpilogue& epi, cudaStream_t stream) {
  const bool a_mn = !ta, b_mn = tb;
  const int total = (m / 256) * n;
  // widest column tile <= 256 that divides n (K-major B halves need 16-row boxes, MN-major 32-column chunks)
  const int gran = b_mn ? 64 : 32;
  int wmax = SPLIT3 ? 128 : 256;  // SPLIT3 keeps hi and lo accumulators of two chunks in the 512 TMEM columns
  while (wmax > gran && n % wmax != 0) wmax -= gran;
  const int tiles = total / wmax;
  const int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
  const CUtensorMap ma = make_operand_map(A, a_mn, m, k, lda, BM);
  const CUtensorMap mb_big = make_operand_map(B, b_mn, n, k, ldb, wmax / 2);
  const CUtensorMap mb_small = make_operand_map(B, b_mn, n, k, ldb, b_mn ? 32 : 16);
  if (a_mn && !b_mn) launch_pair_variant<true, false, SPLIT3>(ma, mb_big, mb_small, C, m, n, k, ldc, epi, wmax, pairs, stream);
  else if (a_mn && b_mn) launch_pair_variant<true, true, SPLIT3>(ma, mb_big, mb_small, C, m, n, k, ldc, epi, wmax, pairs, stream);
  else if (!a_mn && !b_mn) launch_pair_variant<false, false, SPLIT3>(ma, mb_big, mb_small, C, m, n, k, ldc, epi, wmax, pairs, stream);
  else launch_pair_variant<false, true, SPLIT3>(ma, mb_big, mb_small, C, m, n, k, ldc, epi, wmax, pairs, stream);
}

}  // namespace

void gemm_tf32_set_pair(int on) { g_pair_enabled = on; }
void gemm_tf32_set_max_ctas(int n) { g_max_ctas = n < 1 ? 1 : (n > kNumSMs ? kNumSMs : n); }

// A "true dense GEMM" (north_star item 2): whole 128-row / 128-column tiles, K in 32-float swizzle rows, and
// enough work to fill the machine; everything else stays on the FP32 SIMT path.
bool gemm_tf32_eligible(int m, int n, int k, bool, bool) {
  return m >= 256 && n >= 256 && k >= 128 && m % BM == 0 && n % 128 == 0 && k % BK == 0;
}

bool gemm_tf32_aligned(const float* A, const float* B, int lda, int ldb) {
  return ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 127) == 0 && lda % 4 == 0 && ldb % 4 == 0;
}

void launch_gemm_tf32(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb,
                      int ldc, const GemmEpilogue& epi, cudaStream_t stream) {
  if (!gemm_tf32_aligned(A, B, lda, ldb))
    throw Error("tcgen05 product: operands must be 128-byte aligned with leading dimensions in multiples of 4");
  if (pair_eligible(m, n, k)) return launch_pair<false>(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
  if (n % 256 == 0) launch_bn<256>(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
  else launch_bn<128>(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
}

// FP32-accurate product on the tensor core (SPLIT3 above): the CTA-pair kernel only.
bool gemm_tf32x3_eligible(int m, int n, int k, bool ta, bool tb) { return gemm_tf32_eligible(m, n, k, ta, tb) && pair_shape(m, n, k); }

void launch_gemm_tf32x3(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb, int ldc,
                        const GemmEpilogue& epi, cudaStream_t stream) {
  if (!gemm_tf32_aligned(A, B, lda, ldb))
    throw Error("tcgen05 product: operands must be 128-byte aligned with leading dimensions in multiples of 4");
  launch_pair<true>(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
}

}  // namespace cg
