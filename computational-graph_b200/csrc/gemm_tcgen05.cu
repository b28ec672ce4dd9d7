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
constexpr int STAGES = 4;
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

struct TuneParams {
  // descriptor strides; defaults follow the canonical layouts, overridable from the environment for bring-up
  uint32_t k_lbo, k_sbo, k_step;     // K-major operand
  uint32_t mn_lbo, mn_sbo, mn_step;  // MN-major operand
  uint32_t mn_layout;                // descriptor layout type of the MN-major operand
  uint32_t mn_chunked;               // 1: fetch MN-major tiles as 2-D boxes {32, BK}, one per 32-wide chunk
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     float* __restrict__ C, int m, int n, int k, int ldc, int epi_mode, float lr, TuneParams tp) {
  // tp.mn_chunked: the MN-major tile is fetched as ROWS/32 two-dimensional boxes instead of one 3-D box
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
          if (A_MN) {
            if (tp.mn_chunked) {
              for (int c = 0; c < BM / 32; c++) tma_load_2d(sa + c * (BK * 128), &map_a, &full_bar[stage], m0 + c * 32, kb * BK);
            } else {
              tma_load_3d(sa, &map_a, &full_bar[stage], 0, kb * BK, m0 / 32);
            }
          } else {
            tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m0);
          }
          if (B_MN) {
            if (tp.mn_chunked) {
              for (int c = 0; c < BN / 32; c++) tma_load_2d(sb + c * (BK * 128), &map_b, &full_bar[stage], n0 + c * 32, kb * BK);
            } else {
              tma_load_3d(sb, &map_b, &full_bar[stage], 0, kb * BK, n0 / 32);
            }
          } else {
            tma_load_2d(sb, &map_b, &full_bar[stage], kb * BK, n0);
          }
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
            const uint64_t da = A_MN ? make_smem_desc(sa + kk * tp.mn_step, tp.mn_lbo, tp.mn_sbo, tp.mn_layout)
                                     : make_smem_desc(sa + kk * tp.k_step, tp.k_lbo, tp.k_sbo, 2);
            const uint64_t db = B_MN ? make_smem_desc(sb + kk * tp.mn_step, tp.mn_lbo, tp.mn_sbo, tp.mn_layout)
                                     : make_smem_desc(sb + kk * tp.k_step, tp.k_lbo, tp.k_sbo, 2);
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

CUtensorMapSwizzle mn_swizzle() {
  if (const char* e = getenv("CG_TF32_MN_TMA_SWIZZLE")) return (CUtensorMapSwizzle)atoi(e);
  return CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
}

// Tensor map of one operand tile. `mn` = extent along M (A) or N (B); `kdim` = contraction extent;
// `ld` = stored leading dimension (floats); rows = tile extent along mn.
CUtensorMap make_operand_map(const float* base, bool mn_major, int mn, int kdim, int ld, int rows, bool chunked) {
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
  } else if (chunked) {
    cuuint64_t dims[2] = {(cuuint64_t)mn, (cuuint64_t)kdim};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)BK};
    cuuint32_t estr[2] = {1, 1};
    r = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, mn_swizzle(), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // contiguous along MN: element (x, kk) at base[kk * ld + x]; viewed as {32, K, MN/32} so that one box
    // lands as [MN/32 chunks][BK k-rows][32 floats] — the canonical MN-major SWIZZLE_128B layout
    cuuint64_t dims[3] = {32, (cuuint64_t)kdim, (cuuint64_t)(mn / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)(rows / 32)};
    cuuint32_t estr[3] = {1, 1, 1};
    r = g_encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, mn_swizzle(), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return map;
}

TuneParams tune_params() {
  static TuneParams tp = [] {
    TuneParams t;
    t.k_lbo = 16, t.k_sbo = 1024, t.k_step = UMMA_K * 4;          // 8-row atoms 1024 B apart; +32 B per K step
    t.mn_lbo = BK * 128, t.mn_sbo = 512, t.mn_step = 1024;        // 32-wide MN chunks 4096 B apart; 4-row atoms; +8 k-rows per step
    t.mn_layout = 1;
    auto env = [](const char* name, uint32_t& v) {
      if (const char* s = getenv(name)) v = (uint32_t)strtoul(s, nullptr, 0);
    };
    env("CG_TF32_K_LBO", t.k_lbo), env("CG_TF32_K_SBO", t.k_sbo), env("CG_TF32_K_STEP", t.k_step);
    env("CG_TF32_MN_LBO", t.mn_lbo), env("CG_TF32_MN_SBO", t.mn_sbo), env("CG_TF32_MN_STEP", t.mn_step);
    t.mn_chunked = 0;
    env("CG_TF32_MN_CHUNKED", t.mn_chunked), env("CG_TF32_MN_LAYOUT", t.mn_layout);
    return t;
  }();
  return tp;
}

template <int BN, bool A_MN, bool B_MN>
void launch_variant(const CUtensorMap& ma, const CUtensorMap& mb, float* C, int m, int n, int k, int ldc,
                    const GemmEpilogue& epi, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (BM + BN) * BK * 4 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    CG_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int tiles = (m / BM) * (n / BN);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  gemm_tf32_kernel<BN, A_MN, B_MN><<<grid, NUM_THREADS, smem, stream>>>(ma, mb, C, m, n, k, ldc, epi.mode, epi.lr, tune_params());
  CG_CUDA(cudaGetLastError());
}

template <int BN>
void launch_bn(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb, int ldc,
               const GemmEpilogue& epi, cudaStream_t stream) {
  // op(A)(i, kk): not transposed => A[kk * lda + i] (M-contiguous, MN-major); transposed => A[i * lda + kk] (K-major)
  // op(B)(kk, j): not transposed => B[j * ldb + kk] (K-major);                transposed => B[kk * ldb + j] (MN-major)
  const bool a_mn = !ta, b_mn = tb;
  const bool chunked = tune_params().mn_chunked != 0;
  const CUtensorMap ma = make_operand_map(A, a_mn, m, k, lda, BM, chunked);
  const CUtensorMap mb = make_operand_map(B, b_mn, n, k, ldb, BN, chunked);
  if (a_mn && !b_mn) launch_variant<BN, true, false>(ma, mb, C, m, n, k, ldc, epi, stream);
  else if (a_mn && b_mn) launch_variant<BN, true, true>(ma, mb, C, m, n, k, ldc, epi, stream);
  else if (!a_mn && !b_mn) launch_variant<BN, false, false>(ma, mb, C, m, n, k, ldc, epi, stream);
  else launch_variant<BN, false, true>(ma, mb, C, m, n, k, ldc, epi, stream);
}

}  // namespace

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
  if (n % 256 == 0) launch_bn<256>(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
  else launch_bn<128>(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
}

}  // namespace cg
