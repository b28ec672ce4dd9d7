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

#include <mutex>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "cg_common.h"

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
  bool declared[EW_REGS] = {};
  // inputs first: every load is issued before the first dependent instruction
  for (uint32_t k = 0; k < p.n_in; k++) {
    s << "    float r" << k << "[4];\n";
    declared[k] = true;
    if ((p.in_scalar_mask >> k) & 1u) s << "    FOR4 r" << k << "[j] = s" << k << ";\n";
    else s << "    ld4(p.in[" << k << "], b, n, r" << k << ");\n";
  }
  s << "    float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};\n";
  for (uint32_t pc = 0; pc < p.n_ins; pc++) {
    const unsigned ins = p.code[pc], op = ins >> 4, k = ins & 15u;
    auto reg = [&](bool write) {
      if (k >= (unsigned)EW_REGS) throw Error("internal: elementwise register out of range");
      if (!declared[k]) {
        if (!write) throw Error("internal: elementwise register read before write");
        s << "    float r" << k << "[4];\n";
        declared[k] = true;
      }
    };
    switch (op) {
      case OP_LDR: reg(false); s << "    FOR4 a[j] = r" << k << "[j];\n"; break;
      case OP_STR: reg(true); s << "    FOR4 r" << k << "[j] = a[j];\n"; break;
      case OP_LDI: s << "    FOR4 a[j] = p.imm[" << k << "];\n"; break;
      case OP_ADD: reg(false); s << "    FOR4 a[j] = __fadd_rn(a[j], r" << k << "[j]);\n"; break;
      case OP_SUB: reg(false); s << "    FOR4 a[j] = __fsub_rn(a[j], r" << k << "[j]);\n"; break;
      case OP_RSUB: reg(false); s << "    FOR4 a[j] = __fsub_rn(r" << k << "[j], a[j]);\n"; break;
      case OP_MUL: reg(false); s << "    FOR4 a[j] = __fmul_rn(a[j], r" << k << "[j]);\n"; break;
      case OP_MULI: s << "    FOR4 a[j] = __fmul_rn(a[j], p.imm[" << k << "]);\n"; break;
      case OP_NEG: s << "    FOR4 a[j] = -a[j];\n"; break;
      case OP_SQ: s << "    FOR4 a[j] = __fmul_rn(a[j], a[j]);\n"; break;
      case OP_ABS_M: s << "    FOR4 a[j] = a[j] < 0.0f ? -a[j] : a[j];\n"; break;
      case OP_SGN_M: s << "    FOR4 a[j] = a[j] < 0.0f ? -1.0f : 1.0f;\n"; break;
      case OP_ABS_S: s << "    FOR4 a[j] = fabsf(a[j]);\n"; break;
      case OP_SGN_S: s << "    FOR4 a[j] = ew_sgn_scalar(a[j]);\n"; break;
      case OP_SIGMOID: s << "    FOR4 a[j] = ew_sigmoid(a[j]);\n"; break;
      case OP_TANH: s << "    FOR4 a[j] = tanhf(a[j]);\n"; break;
      case OP_ONE_MINUS: s << "    FOR4 a[j] = __fsub_rn(1.0f, a[j]);\n"; break;
      case OP_STG: s << "    st4(p.out[" << k << "], b, n, a);\n"; break;
      default: throw Error("internal: unknown elementwise opcode");
    }
  }
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
    return r;
  }();
  return d;
}

void drv_check(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) throw Error(std::string("CUDA driver error ") + std::to_string((int)r) + " in " + what);
}

std::vector<char> compile_cubin(const std::string& src) {
  Nvrtc& n = nvrtc();
  nvrtcProgram prog = nullptr;
  if (n.CreateProgram(&prog, src.c_str(), "ew_jit.cu", 0, nullptr, nullptr) != 0) throw Error("nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const int rc = n.CompileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) n.GetProgramLog(prog, log.data());
    n.DestroyProgram(&prog);
    throw Error("NVRTC failed to compile a fused elementwise program:\n" + log);
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

}  // namespace cg
