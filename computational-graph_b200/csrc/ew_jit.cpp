// ew_jit.cpp — plan-time specialisation of fused elementwise programs.
//
// The bytecode interpreter (ew_interp.cuh) pays a two-level dispatch per micro-op; on the 22-op C2 program that is
// 2337 instructions per float4 and the kernel ends up instruction-bound at 0.14 of HBM bandwidth
// (profiles/r01_ew_c2_full.raw.txt). For programs that run over large matrices the tape compiler therefore unrolls
// the SAME bytecode into straight-line CUDA C at plan time — one statement per micro-op, the accumulator machine's
// registers become plain float[4] locals — and compiles it for sm_100a with NVRTC. Every arithmetic micro-op is still
// one explicitly rounded IEEE f32 intrinsic (__fadd_rn / __fmul_rn / __fsub_rn / __fdiv_rn, expf, tanhf), so the
// specialised kernel is bit-identical to the interpreter (checked in tests/test_gpu_graph.py) and to the reference's
// one-kernel-per-node evaluation (src/cpu/ops.cpp:68-86). Kernels are cached per program signature; small programs
// and machines without libnvrtc keep the interpreter.
#include <cuda.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>
#include <set>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "tape.h"

namespace cg {

struct EwKernel {
  CUmodule module = nullptr;
  CUfunction fn = nullptr;
  int blocks_per_sm = 4;  // resident CTAs of 256 threads (occupancy query): the grid is one full wave of them
};

namespace {

// ---- NVRTC, resolved at run time (libnvrtc.so.12 ships with the CUDA toolkit of this image) -----------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* lib = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  bool ok = false;
};

Nvrtc& nvrtc() {
  static Nvrtc n = [] {
    Nvrtc r;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      r.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (r.lib) break;
    }
    if (!r.lib) return r;
    auto sym = [&](const char* s) { return dlsym(r.lib, s); };
    r.CreateProgram = reinterpret_cast<decltype(r.CreateProgram)>(sym("nvrtcCreateProgram"));
    r.CompileProgram = reinterpret_cast<decltype(r.CompileProgram)>(sym("nvrtcCompileProgram"));
    r.GetCUBINSize = reinterpret_cast<decltype(r.GetCUBINSize)>(sym("nvrtcGetCUBINSize"));
    r.GetCUBIN = reinterpret_cast<decltype(r.GetCUBIN)>(sym("nvrtcGetCUBIN"));
    r.GetProgramLogSize = reinterpret_cast<decltype(r.GetProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
    r.GetProgramLog = reinterpret_cast<decltype(r.GetProgramLog)>(sym("nvrtcGetProgramLog"));
    r.DestroyProgram = reinterpret_cast<decltype(r.DestroyProgram)>(sym("nvrtcDestroyProgram"));
    r.ok = r.CreateProgram && r.CompileProgram && r.GetCUBINSize && r.GetCUBIN && r.GetProgramLogSize &&
           r.GetProgramLog && r.DestroyProgram;
    return r;
  }();
  return n;
}

// The kernel parameter is the interpreter's EwProgram, byte for byte (cg_common.h); the code[] array is ignored.
const char* kPreamble = R"SRC(
typedef unsigned long long u64;
struct EwProgram {
  u64 n;
  unsigned n_in, n_out, n_ins, in_scalar_mask;
  const float* in[12];
  float* out[12];
  float imm[8];
  unsigned short code[160];
};
#define FOR4 _Pragma("unroll") for (int j = 0; j < 4; j++)
__device__ __forceinline__ void ld4(const float* __restrict__ p, u64 b, u64 n, float (&d)[4]) {
  if (b + 4 <= n) {
    const float4 x = *reinterpret_cast<const float4*>(p + b);
    d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
  } else {
    FOR4 d[j] = (b + j < n) ? p[b + j] : 0.0f;
  }
}
__device__ __forceinline__ void st4(float* __restrict__ p, u64 b, u64 n, const float (&s)[4]) {
  if (b + 4 <= n) {
    *reinterpret_cast<float4*>(p + b) = make_float4(s[0], s[1], s[2], s[3]);
  } else {
    FOR4 if (b + j < n) p[b + j] = s[j];
  }
}
// cpu/ops.cpp:72 — 1.0f / (1.0f + expf(-x)), IEEE division
__device__ __forceinline__ float ew_sigmoid(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
// f32::signum (ops.rs:332): -0.0 -> -1, NaN -> NaN
__device__ __forceinline__ float ew_sgn_scalar(float x) {
  if (x != x) return x;
  return (__float_as_uint(x) >> 31) ? -1.0f : 1.0f;
}
)SRC";

// Operand spelling of one program: where input k / output k live and how immediate k is written. The per-kernel
// specialiser reads them from the EwProgram kernel parameter; the tape specialiser bakes shared-memory offsets and
// literal constants in.
struct EwOperands {
  std::function<std::string(uint32_t)> in, out, imm;
};

// Emits, at the current position of `s`, the straight-line code of program `p` for the float4 at element offset `b`
// of a domain of `n` elements (both are names in scope). Scalar operands are expected in `s<k>`.
void emit_ew_body(std::ostringstream& s, const EwProgram& p, const EwOperands& o, const char* ind, int vec = 4) {
  // vec = 4: one float4 per thread (FOR4 / ld4 / st4); vec = 1: one element per thread (FOR1 / ld1 / st1), for
  // domains so small that spreading them over more threads beats instruction-level parallelism
  const std::string V = std::to_string(vec), FOR = "FOR" + V + " ", LD = "ld" + V + "(", ST = "st" + V + "(";
  bool declared[EW_REGS] = {};
  // inputs first: every load is issued before the first dependent instruction
  for (uint32_t k = 0; k < p.n_in; k++) {
    s << ind << "float r" << k << "[" << V << "];\n";
    declared[k] = true;
    if ((p.in_scalar_mask >> k) & 1u) s << ind << FOR << "r" << k << "[j] = s" << k << ";\n";
    else s << ind << LD << o.in(k) << ", b, n, r" << k << ");\n";
  }
  s << ind << "float a[" << V << "] = {};\n";
  for (uint32_t pc = 0; pc < p.n_ins; pc++) {
    const unsigned ins = p.code[pc], op = ins >> 4, k = ins & 15u;
    auto reg = [&](bool write) {
      if (k >= (unsigned)EW_REGS) throw Error("internal: elementwise register out of range");
      if (!declared[k]) {
        if (!write) throw Error("internal: elementwise register read before write");
        s << ind << "float r" << k << "[" << V << "];\n";
        declared[k] = true;
      }
    };
    s << ind;
    switch (op) {
      case OP_LDR: reg(false); s << FOR << "a[j] = r" << k << "[j];\n"; break;
      case OP_STR: reg(true); s << FOR << "r" << k << "[j] = a[j];\n"; break;
      case OP_LDI: s << FOR << "a[j] = " << o.imm(k) << ";\n"; break;
      case OP_ADD: reg(false); s << FOR << "a[j] = __fadd_rn(a[j], r" << k << "[j]);\n"; break;
      case OP_SUB: reg(false); s << FOR << "a[j] = __fsub_rn(a[j], r" << k << "[j]);\n"; break;
      case OP_RSUB: reg(false); s << FOR << "a[j] = __fsub_rn(r" << k << "[j], a[j]);\n"; break;
      case OP_MUL: reg(false); s << FOR << "a[j] = __fmul_rn(a[j], r" << k << "[j]);\n"; break;
      case OP_MULI: s << FOR << "a[j] = __fmul_rn(a[j], " << o.imm(k) << ");\n"; break;
      case OP_NEG: s << FOR << "a[j] = -a[j];\n"; break;
      case OP_SQ: s << FOR << "a[j] = __fmul_rn(a[j], a[j]);\n"; break;
      case OP_ABS_M: s << FOR << "a[j] = a[j] < 0.0f ? -a[j] : a[j];\n"; break;
      case OP_SGN_M: s << FOR << "a[j] = a[j] < 0.0f ? -1.0f : 1.0f;\n"; break;
      case OP_ABS_S: s << FOR << "a[j] = fabsf(a[j]);\n"; break;
      case OP_SGN_S: s << FOR << "a[j] = ew_sgn_scalar(a[j]);\n"; break;
      case OP_SIGMOID: s << FOR << "a[j] = ew_sigmoid(a[j]);\n"; break;
      case OP_TANH: s << FOR << "a[j] = tanhf(a[j]);\n"; break;
      case OP_ONE_MINUS: s << FOR << "a[j] = __fsub_rn(1.0f, a[j]);\n"; break;
      case OP_STG: s << ST << o.out(k) << ", b, n, a);\n"; break;
      default: throw Error("internal: unknown elementwise opcode");
    }
  }
}

std::string generate_source(const EwProgram& p) {
  std::ostringstream s;
  s << kPreamble;
  s << "extern \"C\" __global__ void __launch_bounds__(256) ew_jit(const __grid_constant__ EwProgram p) {\n"
       "  const u64 n = p.n;\n"
       "  const u64 step = (u64)gridDim.x * 1024;\n";
  // scalar (broadcast) operands are loop invariant
  for (uint32_t k = 0; k < p.n_in; k++)
    if ((p.in_scalar_mask >> k) & 1u) s << "  const float s" << k << " = *p.in[" << k << "];\n";
  s << "  for (u64 b = ((u64)blockIdx.x * 256 + threadIdx.x) * 4; b < n; b += step) {\n";
  EwOperands o;
  o.in = [](uint32_t k) { return "p.in[" + std::to_string(k) + "]"; };
  o.out = [](uint32_t k) { return "p.out[" + std::to_string(k) + "]"; };
  o.imm = [](uint32_t k) { return "p.imm[" + std::to_string(k) + "]"; };
  emit_ew_body(s, p, o, "    ");
  s << "  }\n}\n";
  return s.str();
}

std::string signature(const EwProgram& p) {
  std::string key;
  key.reserve(16 + p.n_ins * 2);
  auto put = [&](uint32_t v) { key.append(reinterpret_cast<const char*>(&v), sizeof v); };
  put(p.n_in), put(p.n_out), put(p.n_ins), put(p.in_scalar_mask);
  key.append(reinterpret_cast<const char*>(p.code), p.n_ins * sizeof(uint16_t));
  return key;
}

std::mutex g_mu;
std::unordered_map<std::string, EwKernel*> g_cache;  // nullptr = compilation failed once: do not retry
bool g_warned = false;

// Driver entry points, resolved through the runtime so that libcg_b200.so carries no link-time dependency on
// libcuda (the library must load on a CPU-only box for the ABI tests).
struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**) = nullptr;
  CUresult (*Occupancy)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
};

Driver& driver() {
  static Driver d = [] {
    Driver r;
    auto get = [](const char* name) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        throw Error(std::string("driver entry point not available: ") + name);
      return fn;
    };
    r.ModuleLoadData = reinterpret_cast<decltype(r.ModuleLoadData)>(get("cuModuleLoadData"));
    r.ModuleGetFunction = reinterpret_cast<decltype(r.ModuleGetFunction)>(get("cuModuleGetFunction"));
    r.LaunchKernel = reinterpret_cast<decltype(r.LaunchKernel)>(get("cuLaunchKernel"));
    r.Occupancy = reinterpret_cast<decltype(r.Occupancy)>(get("cuOccupancyMaxActiveBlocksPerMultiprocessor"));
    r.FuncSetAttribute = reinterpret_cast<decltype(r.FuncSetAttribute)>(get("cuFuncSetAttribute"));
    r.ModuleUnload = reinterpret_cast<decltype(r.ModuleUnload)>(get("cuModuleUnload"));
    return r;
  }();
  return d;
}

void drv_check(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) throw Error(std::string("CUDA driver error ") + std::to_string((int)r) + " in " + what);
}

std::vector<char> compile_cubin(const std::string& src, const char* name = "ew_jit.cu") {
  Nvrtc& n = nvrtc();
  nvrtcProgram prog = nullptr;
  if (n.CreateProgram(&prog, src.c_str(), name, 0, nullptr, nullptr) != 0) throw Error("nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const int rc = n.CompileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) n.GetProgramLog(prog, log.data());
    n.DestroyProgram(&prog);
    throw Error(std::string("NVRTC failed to compile ") + name + ":\n" + log);
  }
  size_t sz = 0;
  n.GetCUBINSize(prog, &sz);
  std::vector<char> cubin(sz);
  n.GetCUBIN(prog, cubin.data());
  n.DestroyProgram(&prog);
  return cubin;
}

}  // namespace

bool ew_jit_available() { return nvrtc().ok; }

// Source of the specialised kernel (also used by the CPU-side unit test, which compiles it without a GPU).
std::string ew_jit_source(const EwProgram& p) { return generate_source(p); }

size_t ew_jit_compile_only(const EwProgram& p) {
  if (!nvrtc().ok) throw Error("libnvrtc is not available");
  return compile_cubin(generate_source(p)).size();
}

const EwKernel* ew_jit_get(const EwProgram& p) {
  if (!nvrtc().ok) {
    if (!g_warned) {
      fprintf(stderr, "libcg_b200: libnvrtc.so.12 not found — fused elementwise programs use the interpreter kernel\n");
      g_warned = true;
    }
    return nullptr;
  }
  const std::string key = signature(p);
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) return it->second;
  const std::vector<char> cubin = compile_cubin(generate_source(p));
  auto* k = new EwKernel();
  CG_CUDA(cudaFree(nullptr));  // make sure the primary context is current on this thread
  drv_check(driver().ModuleLoadData(&k->module, cubin.data()), "cuModuleLoadData");
  drv_check(driver().ModuleGetFunction(&k->fn, k->module, "ew_jit"), "cuModuleGetFunction");
  int occ = 0;
  if (driver().Occupancy(&occ, k->fn, 256, 0) == CUDA_SUCCESS && occ > 0) k->blocks_per_sm = occ;
  g_cache[key] = k;
  return k;
}

size_t ew_jit_cache_size() {
  std::lock_guard<std::mutex> lock(g_mu);
  return g_cache.size();
}

void launch_ew_jit(const EwKernel* k, const EwProgram& p, cudaStream_t stream) {
  if (p.n == 0) return;
  const size_t tiles = (p.n + 1023) / 1024;
  // grid-stride loop over exactly one resident wave (a ragged second wave costs up to a third of the kernel)
  const size_t wave = (size_t)kNumSMs * (size_t)k->blocks_per_sm;
  const size_t grid = tiles < wave ? tiles : wave;
  void* params[] = {const_cast<EwProgram*>(&p)};
  drv_check(driver().LaunchKernel(k->fn, (unsigned)grid, 1, 1, 256, 1, 1, 0, (CUstream)stream, params, nullptr), "cuLaunchKernel");
}

// =====================================================================================================================
// Whole-tape specialisation: the resident executor of vm_kernel.cu, compiled.
//
// For a latency-bound plan (BASELINE config 3: the README LSTM, every matrix <= 30 x 30) the interpreter VM still
// pays a dependent chain of global loads per step — step record, program, bytecode, operands out of L2 — about 1 us a
// step. Here the plan is unrolled at plan time into ONE straight-line kernel: every step is inline code with its
// shapes as compile-time constants, and EVERY VALUE OF THE PLAN — leaves, arena temporaries, the root — lives in the
// CTA's shared memory for the whole `minimize` call (the README LSTM at d = 30 needs ~180 KB of the SM's 227 KB).
// Leaves are loaded once when the kernel starts and written back once when it ends; in between an iteration touches
// neither L2 nor HBM. Products are 2 x 2 register-blocked over shared-memory operands, every output still accumulated
// k-ascending with separate multiply / add roundings (cpu/ops.cpp:103-114), so the bits are the per-step kernels'.
// =====================================================================================================================
// The compiled code depends only on the plan's STRUCTURE (shapes, programs, immediates, shared-memory layout): the
// addresses of the persistent buffers are a kernel parameter, so structurally identical plans — the reference's test
// suite builds the same few graphs over and over — share one module.
struct TapeModule {
  CUmodule module = nullptr;
  CUfunction fn = nullptr;
  size_t cubin_bytes = 0;
};
struct TapeKernel {
  const TapeModule* mod = nullptr;
  size_t smem_bytes = 0;
  std::vector<float*> ptrs;  // persistent buffers, in the order of the kernel's pointer table
};

namespace {

constexpr int TAPE_THREADS = 512;
constexpr size_t TAPE_SMEM_BUDGET = 224 * 1024;  // of the 227 KB a CTA may own on sm_100a
// Domains up to this many elements get one element (one product output) per thread; larger ones a float4 / a 2 x 2
// block. Measured on the README LSTM (exact mode): d = 15 (225 elements) 20.0 vs 26.3 us per iteration in favour of
// one per thread, d = 20 (400 elements) 41.2 vs 36.3 us in favour of blocks.
constexpr size_t TAPE_SMALL = 256;

const char* kTapePreamble = R"SRC(
typedef unsigned long long u64;
#define FOR4 _Pragma("unroll") for (int j = 0; j < 4; j++)
#define T 512
extern __shared__ __align__(16) float S[];
__device__ __forceinline__ void ld4(const float* p, u64 b, u64 n, float (&d)[4]) {
  if (b + 4 <= n) {
    const float4 x = *reinterpret_cast<const float4*>(p + b);
    d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
  } else {
    FOR4 d[j] = (b + j < n) ? p[b + j] : 0.0f;
  }
}
__device__ __forceinline__ void st4(float* p, u64 b, u64 n, const float (&s)[4]) {
  if (b + 4 <= n) {
    *reinterpret_cast<float4*>(p + b) = make_float4(s[0], s[1], s[2], s[3]);
  } else {
    FOR4 if (b + j < n) p[b + j] = s[j];
  }
}
#define FOR1 for (int j = 0; j < 1; j++)
__device__ __forceinline__ void ld1(const float* p, u64 b, u64, float (&d)[1]) { d[0] = p[b]; }
__device__ __forceinline__ void st1(float* p, u64 b, u64, const float (&s)[1]) { p[b] = s[0]; }
__device__ __forceinline__ float ew_sigmoid(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
__device__ __forceinline__ float ew_sgn_scalar(float x) {
  if (x != x) return x;
  return (__float_as_uint(x) >> 31) ? -1.0f : 1.0f;
}
__device__ __forceinline__ void cp_in(int off, const float* g, int n) {
  for (int i = threadIdx.x; i < n; i += T) S[off + i] = g[i];
}
__device__ __forceinline__ void cp_out(float* g, int off, int n) {
  for (int i = threadIdx.x; i < n; i += T) g[i] = S[off + i];
}
template <int EPI>
__device__ __forceinline__ void put(int co, int ldc, int i, int j, float acc, float lr) {
  float* c = S + co + j * ldc + i;
  if (EPI == 1) *c = __fsub_rn(*c, __fmul_rn(acc, lr));  // optimize.rs:149-152: two roundings
  else if (EPI == 2) *c = __fadd_rn(*c, acc);
  else *c = acc;
}
// C(M x N) = op(A) . op(B) over shared memory, column-major, 2 x 2 outputs per thread; each output is the k-ascending
// separate-rounding sum of the reference (cpu/ops.cpp:103-114). One instantiation per distinct product shape.
template <bool TA, bool TB, int EPI, int M, int N, int K, int LDA, int LDB, int LDC>
__device__ __noinline__ void gemm(int ao, int bo, int co, float lr, int shift) {
  constexpr int MB = (M + 1) / 2, NB = (N + 1) / 2;
  for (int o = (threadIdx.x + T - shift) & (T - 1); o < MB * NB; o += T) {
    const int i0 = (o % MB) * 2, j0 = (o / MB) * 2;
    const bool i1 = i0 + 1 < M, j1 = j0 + 1 < N;
    const int ia = i1 ? i0 + 1 : i0, jb = j1 ? j0 + 1 : j0;  // clamped: the duplicate lane is computed, never stored
    float c00 = 0.0f, c01 = 0.0f, c10 = 0.0f, c11 = 0.0f;
    // adjacent operands come in one 8-byte load where the layout makes them adjacent and aligned (offsets are multiples
    // of 4 floats, i0 / j0 are even): rows i0, i0+1 of a non-transposed A, columns j0, j0+1 of a transposed B
    constexpr bool A2 = !TA && LDA % 2 == 0, B2 = TB && LDB % 2 == 0;
#pragma unroll 5
    for (int kk = 0; kk < K; kk++) {
      float a0, a1, b0, b1;
      if (A2 && i1) {
        const float2 v = *reinterpret_cast<const float2*>(&S[ao + kk * LDA + i0]);
        a0 = v.x, a1 = v.y;
      } else {
        a0 = S[ao + (TA ? i0 * LDA + kk : kk * LDA + i0)];
        a1 = S[ao + (TA ? ia * LDA + kk : kk * LDA + ia)];
      }
      if (B2 && j1) {
        const float2 v = *reinterpret_cast<const float2*>(&S[bo + kk * LDB + j0]);
        b0 = v.x, b1 = v.y;
      } else {
        b0 = S[bo + (TB ? kk * LDB + j0 : j0 * LDB + kk)];
        b1 = S[bo + (TB ? kk * LDB + jb : jb * LDB + kk)];
      }
      c00 = __fadd_rn(c00, __fmul_rn(a0, b0));
      c01 = __fadd_rn(c01, __fmul_rn(a0, b1));
      c10 = __fadd_rn(c10, __fmul_rn(a1, b0));
      c11 = __fadd_rn(c11, __fmul_rn(a1, b1));
    }
    put<EPI>(co, LDC, i0, j0, c00, lr);
    if (j1) put<EPI>(co, LDC, i0, j0 + 1, c01, lr);
    if (i1) put<EPI>(co, LDC, i0 + 1, j0, c10, lr);
    if (i1 && j1) put<EPI>(co, LDC, i0 + 1, j0 + 1, c11, lr);
  }
}
// The same product with one output per thread: for small shapes (M * N <= TAPE_SMALL) every output gets its own thread and
// the per-thread instruction count — what bounds a one-warp step — is a quarter of the blocked version's.
template <bool TA, bool TB, int EPI, int M, int N, int K, int LDA, int LDB, int LDC>
__device__ __noinline__ void gemm1(int ao, int bo, int co, float lr, int shift) {
  for (int o = (threadIdx.x + T - shift) & (T - 1); o < M * N; o += T) {
    const int i = o % M, j = o / M;
    float acc = 0.0f;
#pragma unroll 6
    for (int kk = 0; kk < K; kk++)
      acc = __fadd_rn(acc, __fmul_rn(S[ao + (TA ? i * LDA + kk : kk * LDA + i)], S[bo + (TB ? kk * LDB + j : j * LDB + kk)]));
    put<EPI>(co, LDC, i, j, acc, lr);
  }
}
)SRC";

struct TapeBuf {
  const float* ptr;
  size_t n, off;
  bool persistent, written;
};

std::string f32_literal(float v) {
  uint32_t bits;
  memcpy(&bits, &v, sizeof bits);
  char buf[48];
  snprintf(buf, sizeof buf, "__int_as_float(0x%08x)", bits);
  return buf;
}

// Lays every buffer the plan touches out in shared memory; false when the plan does not fit (or buffers overlap
// in a way one flat layout cannot express).
bool tape_layout(const Plan& p, std::vector<TapeBuf>& bufs, size_t& total_floats) {
  std::unordered_map<const float*, size_t> index;
  auto touch = [&](const float* ptr, size_t n, bool write) {
    auto it = index.find(ptr);
    if (it == index.end()) {
      index[ptr] = bufs.size();
      bufs.push_back(TapeBuf{ptr, n, 0, false, write});
    } else {
      TapeBuf& b = bufs[it->second];
      b.n = std::max(b.n, n);
      b.written |= write;
    }
  };
  for (const Step& s : p.steps) {
    if (s.kind == Step::EW) {
      for (uint32_t k = 0; k < s.prog.n_in; k++) touch(s.prog.in[k], ((s.prog.in_scalar_mask >> k) & 1u) ? 1 : s.prog.n, false);
      for (uint32_t k = 0; k < s.prog.n_out; k++) touch(s.prog.out[k], s.prog.n, true);
    } else if (s.kind == Step::AVG) {
      touch(p.vals[s.ins.a].ptr, p.vals[s.ins.a].sh.n(), false);
      touch(p.vals[s.ins.dst].ptr, 1, true);
    } else if (s.kind == Step::GEMM) {
      touch(p.vals[s.ins.a].ptr, p.vals[s.ins.a].sh.n(), false);
      touch(p.vals[s.ins.b].ptr, p.vals[s.ins.b].sh.n(), false);
      touch(p.vals[s.ins.dst].ptr, p.vals[s.ins.dst].sh.n(), true);
    } else {
      return false;
    }
  }
  touch(p.root_ptr, p.root_n, false);
  // no partial overlaps (arena slots are recycled whole, leaves are separate allocations)
  std::vector<const TapeBuf*> sorted;
  for (const TapeBuf& b : bufs) sorted.push_back(&b);
  std::sort(sorted.begin(), sorted.end(), [](const TapeBuf* a, const TapeBuf* b) { return a->ptr < b->ptr; });
  for (size_t i = 0; i + 1 < sorted.size(); i++)
    if (sorted[i]->ptr + sorted[i]->n > sorted[i + 1]->ptr) return false;
  const float* alo = p.arena ? p.arena->ptr : nullptr;
  const float* ahi = p.arena ? p.arena->ptr + p.arena->n : nullptr;
  total_floats = 0;
  for (TapeBuf& b : bufs) {
    b.persistent = !(alo && b.ptr >= alo && b.ptr < ahi);
    b.off = total_floats;
    total_floats += (b.n + 3) / 4 * 4;
  }
  return total_floats * sizeof(float) <= TAPE_SMEM_BUDGET;
}

std::string tape_source(const Plan& p, const std::vector<TapeBuf>& bufs, std::vector<float*>& table) {
  std::unordered_map<const float*, const TapeBuf*> at;
  for (const TapeBuf& b : bufs) at[b.ptr] = &b;
  auto off = [&](const float* ptr) { return std::to_string(at.at(ptr)->off); };
  std::unordered_map<const float*, size_t> slot;  // persistent buffer -> index in the pointer table
  for (const TapeBuf& b : bufs)
    if (b.persistent) {
      slot[b.ptr] = table.size();
      table.push_back(const_cast<float*>(b.ptr));
    }
  std::ostringstream s;
  s << kTapePreamble;
  s << "struct Ptrs { float* p[" << std::max<size_t>(table.size(), 1) << "]; };\n";
  s << "extern \"C\" __global__ void __launch_bounds__(512, 1) tape_jit(int iters, float* trace, const __grid_constant__ Ptrs P) {\n"
       "  const unsigned t = threadIdx.x;\n";
  for (const TapeBuf& b : bufs)
    if (b.persistent) s << "  cp_in(" << b.off << ", P.p[" << slot[b.ptr] << "], " << b.n << ");\n";
  s << "  __syncthreads();\n"
       "  for (int it = 0; it < iters; it++) {\n";
  const std::string trace_copy = "    if (trace) {\n      cp_out(trace + (u64)it * " + std::to_string(p.root_n) + ", " + off(p.root_ptr) + ", " +
                                 std::to_string(p.root_n) + ");\n      __syncthreads();\n    }\n";
  // the traced value of an iteration is its forward value, before the update (Q12): a leaf root is read first
  if (p.root_is_leaf) s << trace_copy;
  // Steps are scheduled by dependency LEVEL (RAW / WAR / WAW on the buffers they touch, in tape order): a level is one
  // phase between two __syncthreads(), so the barrier count is the critical path of the iteration, not its step count.
  // Independent steps of one phase start at different threads (`shift`), so e.g. the gate products of an LSTM step run
  // side by side instead of one after the other.
  struct Access {
    std::vector<const float*> rd, wr;
    size_t items = 1;
  };
  const size_t nsteps = p.steps.size();
  std::vector<Access> acc(nsteps);
  for (size_t i = 0; i < nsteps; i++) {
    const Step& st = p.steps[i];
    Access& a = acc[i];
    if (st.kind == Step::EW) {
      for (uint32_t k = 0; k < st.prog.n_in; k++) a.rd.push_back(st.prog.in[k]);
      for (uint32_t k = 0; k < st.prog.n_out; k++) a.wr.push_back(st.prog.out[k]);
      a.items = st.prog.n <= TAPE_SMALL ? st.prog.n : (st.prog.n + 3) / 4;
    } else if (st.kind == Step::AVG) {
      a.rd.push_back(p.vals[st.ins.a].ptr);
      a.wr.push_back(p.vals[st.ins.dst].ptr);
    } else {
      a.rd.push_back(p.vals[st.ins.a].ptr);
      a.rd.push_back(p.vals[st.ins.b].ptr);
      if (st.ins.epi) a.rd.push_back(p.vals[st.ins.dst].ptr);
      a.wr.push_back(p.vals[st.ins.dst].ptr);
      const Shape& cs = p.vals[st.ins.dst].sh;
      a.items = cs.n() <= TAPE_SMALL ? cs.n() : (size_t)((cs.h + 1) / 2) * ((cs.w + 1) / 2);
    }
  }
  std::vector<int> level(nsteps, 0);
  {
    struct Loc {
      int last_writer = -1;
      std::vector<int> readers;
    };
    std::unordered_map<const float*, Loc> loc;
    for (size_t i = 0; i < nsteps; i++) {
      int lv = 0;
      auto dep = [&](int d) {
        if (d >= 0) lv = std::max(lv, level[d] + 1);
      };
      for (const float* r : acc[i].rd) dep(loc[r].last_writer);
      for (const float* w : acc[i].wr) {
        Loc& l = loc[w];
        dep(l.last_writer);
        for (int rd : l.readers) dep(rd);
      }
      level[i] = lv;
      for (const float* r : acc[i].rd) loc[r].readers.push_back((int)i);
      for (const float* w : acc[i].wr) {
        Loc& l = loc[w];
        l.last_writer = (int)i;
        l.readers.clear();
      }
    }
  }
  std::vector<size_t> order(nsteps);
  for (size_t i = 0; i < nsteps; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return level[x] < level[y]; });
  unsigned shift = 0;
  int cur_level = 0;
  for (size_t oi = 0; oi < nsteps; oi++) {
    const size_t i = order[oi];
    const Step& st = p.steps[i];
    const size_t items = acc[i].items;
    if (level[i] != cur_level) {
      s << "    __syncthreads();\n";
      cur_level = level[i];
      shift = 0;
    }
    if (st.kind == Step::EW) {
      const EwProgram& pr = st.prog;
      s << "    {  // step " << i << ": elementwise, n = " << pr.n << "\n      const u64 n = " << pr.n << ";\n";
      const int vec = pr.n <= TAPE_SMALL ? 1 : 4;
      s << "      for (u64 b = (u64)((t + T - " << shift << ") & (T - 1)) * " << vec << "; b < n; b += T * " << vec << ") {\n";
      // scalar operands are read INSIDE the loop, i.e. only by the threads that take part in the step: a program that
      // updates a scalar in place (x <- x - e) would otherwise have every idle thread read x while one thread writes
      // it — a value nobody uses, but a data race all the same (found by the ThreadSanitizer run of tests/tape_emu)
      for (uint32_t k = 0; k < pr.n_in; k++)
        if ((pr.in_scalar_mask >> k) & 1u) s << "        const float s" << k << " = S[" << off(pr.in[k]) << "];\n";
      EwOperands o;
      o.in = [&](uint32_t k) { return "S + " + off(pr.in[k]); };
      o.out = [&](uint32_t k) { return "S + " + off(pr.out[k]); };
      o.imm = [&](uint32_t k) { return f32_literal(pr.imm[k]); };
      emit_ew_body(s, pr, o, "        ", vec);
      s << "      }\n    }\n";
    } else if (st.kind == Step::AVG) {
      const Val& a = p.vals[st.ins.a];
      s << "    if (t == " << shift << ") {  // step " << i << ": scalar <- matrix mean, sequential sum (cpu/ops.cpp:138-145)\n"
        << "      float sum = 0.0f;\n      for (int i = 0; i < " << a.sh.n() << "; i++) sum = __fadd_rn(sum, S[" << off(a.ptr) << " + i]);\n"
        << "      S[" << off(p.vals[st.ins.dst].ptr) << "] = __fdiv_rn(sum, " << f32_literal((float)a.sh.n()) << ");\n    }\n";
    } else {
      const Val &a = p.vals[st.ins.a], &b = p.vals[st.ins.b], &c = p.vals[st.ins.dst];
      const int m = (int)c.sh.h, n = (int)c.sh.w, k = (int)(st.ins.ta ? a.sh.h : a.sh.w);
      if (st.ins.epi && p.vals[st.ins.c_in].ptr != c.ptr) throw Error("internal: gemm epilogue operand is not in place");
      s << "    " << (c.sh.n() <= TAPE_SMALL ? "gemm1<" : "gemm<") << (st.ins.ta ? "true" : "false") << ", " << (st.ins.tb ? "true" : "false") << ", " << st.ins.epi << ", " << m
        << ", " << n << ", " << k << ", " << a.sh.h << ", " << b.sh.h << ", " << c.sh.h << ">(" << off(a.ptr) << ", " << off(b.ptr)
        << ", " << off(c.ptr) << ", " << f32_literal(st.ins.imm) << ", " << shift << ");  // step " << i << "\n";
    }
    shift = (unsigned)(((shift + items + 31) / 32 * 32) % TAPE_THREADS);  // the next step of this phase starts on a fresh warp
  }
  s << "    __syncthreads();\n";
  if (!p.root_is_leaf) s << trace_copy;
  s << "  }\n";
  for (const TapeBuf& b : bufs)
    if (b.persistent && b.written) s << "  cp_out(P.p[" << slot[b.ptr] << "], " << b.off << ", " << b.n << ");\n";
  s << "}\n";
  return s.str();
}

}  // namespace

std::unordered_map<std::string, TapeModule*> g_tape_cache;  // source -> module (guarded by g_mu)

const TapeKernel* tape_jit_build(const Plan& p, bool load) {
  if (!nvrtc().ok || p.steps.empty()) return nullptr;
  std::vector<TapeBuf> bufs;
  size_t total = 0;
  if (!tape_layout(p, bufs, total)) return nullptr;
  auto* k = new TapeKernel();
  const std::string src = tape_source(p, bufs, k->ptrs);
  if (k->ptrs.size() > 448) {  // the pointer table is a kernel parameter (4 KB of them at most)
    delete k;
    return nullptr;
  }
  if (const char* path = getenv("CG_TAPE_JIT_DUMP")) {
    if (FILE* f = fopen(path, "w")) {
      fputs(src.c_str(), f);
      fclose(f);
    }
    // Plan-only mode: next to the source, what a launch needs — the pointer table with the initial contents of every
    // persistent buffer — so that tests/tape_emu can execute the generated source on the CPU (test infrastructure;
    // the product never does). Little-endian: u64 count, root_off, root_n, root_is_leaf, smem_floats; then per
    // table entry u64 address, n, offset, written and n floats.
    if (plan_only()) {
      if (FILE* f = fopen((std::string(path) + ".table").c_str(), "wb")) {
        auto put = [&](unsigned long long v) { fwrite(&v, sizeof v, 1, f); };
        std::unordered_map<const float*, const TapeBuf*> at;
        for (const TapeBuf& b : bufs) at[b.ptr] = &b;
        put(k->ptrs.size()), put(at.at(p.root_ptr)->off), put(p.root_n), put(p.root_is_leaf ? 1 : 0), put(total);
        for (float* ptr : k->ptrs) {
          const TapeBuf& b = *at.at(ptr);
          put((unsigned long long)(uintptr_t)ptr), put(b.n), put(b.off), put(b.written ? 1 : 0);
          std::vector<float> init(b.n, 0.0f);
          if (const std::vector<float>* sh = plan_only_shadow(ptr))
            for (size_t i = 0; i < b.n && i < sh->size(); i++) init[i] = (*sh)[i];
          fwrite(init.data(), sizeof(float), b.n, f);
        }
        fclose(f);
      }
    }
  }
  k->smem_bytes = total * sizeof(float);
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_tape_cache.find(src);
  if (it == g_tape_cache.end()) {
    std::vector<char> cubin;
    try {
      cubin = compile_cubin(src, "tape_jit.cu");
    } catch (...) {
      delete k;
      throw;
    }
    auto* m = new TapeModule();
    m->cubin_bytes = cubin.size();
    if (load) {
      CG_CUDA(cudaFree(nullptr));  // make sure the primary context is current on this thread
      drv_check(driver().ModuleLoadData(&m->module, cubin.data()), "cuModuleLoadData");
      drv_check(driver().ModuleGetFunction(&m->fn, m->module, "tape_jit"), "cuModuleGetFunction");
    }
    it = g_tape_cache.emplace(src, m).first;
  }
  k->mod = it->second;
  // (the attribute is per function; the largest request made so far stays in force)
  if (load && k->smem_bytes > 48 * 1024)
    drv_check(driver().FuncSetAttribute(k->mod->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)k->smem_bytes), "cuFuncSetAttribute");
  return k;
}

void tape_jit_free(const TapeKernel* k) { delete k; }  // modules stay cached for the life of the process

size_t tape_jit_cache_size() {
  std::lock_guard<std::mutex> lock(g_mu);
  return g_tape_cache.size();
}

size_t tape_jit_smem_bytes(const TapeKernel* k) { return k ? k->smem_bytes : 0; }

void launch_tape_jit(const TapeKernel* k, int iters, float* trace, cudaStream_t stream) {
  if (iters <= 0) return;
  std::vector<float*> table = k->ptrs;
  if (table.empty()) table.push_back(nullptr);
  void* params[] = {&iters, &trace, table.data()};  // the third parameter is the pointer table, passed by value
  drv_check(driver().LaunchKernel(k->mod->fn, 1, 1, 1, TAPE_THREADS, 1, 1, (unsigned)k->smem_bytes, (CUstream)stream, params, nullptr),
            "cuLaunchKernel");
}

}  // namespace cg
