// dp.cpp — data-parallel plumbing: one process per GPU, one NCCL communicator over the
// NVLink 5 / NVSwitch domain, a SUM all-reduce of every weight gradient before its update (the
// reference has no batch normalisation of the loss — root error == 1, optimize.rs:69 — so gradients
// add, they are not averaged). The reference itself has no distributed code at all.
//
// The sum itself is done by the fused exchange kernel of dp_kernels.cu whenever the ranks can address each other's
// memory: this file opens every gradient slot and weight of a plan in every peer through CUDA IPC ("windows":
// per-buffer (allocation handle, offset) pairs exchanged over an NCCL all-gather, every rank agreeing on the outcome)
// and hands the kernel the peer pointers and the flag blocks. ncclAllReduce remains for weights too small to be
// worth a window, for machines without peer access and for CG_DP_FUSED=0.
//
// NCCL is resolved with dlopen at cg_dp_init time, so the library loads (and every single-GPU path
// runs) on machines without NCCL, and a process that already carries torch's NCCL reuses that copy.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <utility>

#include "../../include/cg_graph.h"
#include "tape.h"

namespace cg {

namespace {
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
Nccl g_nccl;

void load_nccl() {
  if (g_nccl.handle) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) throw Error(std::string("cannot load NCCL: ") + dlerror());
#define NCCL_SYM(field, name)                                            \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);                 \
  if (!g_nccl.field) throw Error(std::string("NCCL symbol missing: ") + name);
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  NCCL_SYM(CommInitRank, "ncclCommInitRank")
  NCCL_SYM(AllReduce, "ncclAllReduce")
  NCCL_SYM(AllGather, "ncclAllGather")
  NCCL_SYM(CommDestroy, "ncclCommDestroy")
  NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
}
void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw Error(std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
}
// ---- peer windows: buffers every rank has opened through CUDA IPC ----------------------------------------------
// An IPC handle names a whole cudaMalloc allocation (small buffers are carved out of a shared block, and the carving
// need not agree between ranks), so every rank publishes, per buffer, (allocation handle, offset inside it); a peer
// opens each allocation once and addresses the buffer as base + that rank's offset.
struct Opened {
  int rank;
  unsigned long long serial;  // the owner's export number: a freed-and-reallocated block gets a new one
  cudaIpcMemHandle_t handle;
  char* base;
};
struct PeerMap {
  const void* local;
  void* peer[DP_MAX_WORLD];
};
std::vector<Opened> g_opened;
std::vector<PeerMap> g_ptrs;
std::vector<std::pair<char*, unsigned long long>> g_exported;  // this rank's allocations that peers may have open
unsigned long long g_next_serial = 1;
uint32_t* g_flags = nullptr;  // this rank's flag block
int g_fused = -1;             // -1 unknown, 0 unavailable (stay on ncclAllReduce), 1 available

void close_windows() {
  for (Opened& o : g_opened) cudaIpcCloseMemHandle(o.base);
  g_opened.clear();
  g_ptrs.clear();
  g_exported.clear();
  if (g_flags) cudaFree(g_flags);
  g_flags = nullptr;
  g_fused = -1;
}

PeerMap* find_ptr(const void* p) {
  for (PeerMap& m : g_ptrs)
    if (m.local == p) return &m;
  return nullptr;
}

// all-gather of `bytes` host bytes per rank over the communicator
std::vector<unsigned char> allgather_bytes(const void* mine, size_t bytes) {
  Runtime& rt = Runtime::get();
  unsigned char *send = nullptr, *recv = nullptr;
  CG_CUDA(cudaMalloc(&send, bytes));
  CG_CUDA(cudaMalloc(&recv, bytes * g_nccl.world));
  CG_CUDA(cudaMemcpyAsync(send, mine, bytes, cudaMemcpyHostToDevice, rt.stream));
  nccl_check(g_nccl.AllGather(send, recv, bytes, ncclUint8, g_nccl.comm, rt.stream), "ncclAllGather");
  std::vector<unsigned char> out(bytes * g_nccl.world);
  CG_CUDA(cudaMemcpyAsync(out.data(), recv, out.size(), cudaMemcpyDeviceToHost, rt.stream));
  CG_CUDA(cudaStreamSynchronize(rt.stream));
  cudaFree(send);
  cudaFree(recv);
  return out;
}

struct AllocMsg {
  cudaIpcMemHandle_t handle;
  unsigned long long serial;
  int ok;
};
struct PtrMsg {
  unsigned long long alloc, offset;
};

// The cudaMalloc allocation that contains `p`.
char* allocation_base(const void* p) {
  typedef CUresult (*Fn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static Fn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return reinterpret_cast<Fn>(f);
  }();
  CUdeviceptr base = 0;
  size_t size = 0;
  if (!fn || fn(&base, &size, (CUdeviceptr)p) != CUDA_SUCCESS) return nullptr;
  return reinterpret_cast<char*>(base);
}

}  // namespace

// plan-only mode (graph.cpp): a pretended world, so that the data-parallel tape can be inspected without GPUs
static int g_plan_world = 1, g_plan_fused = 0;

bool dp_enabled() { return plan_only() ? g_plan_world > 1 : g_nccl.comm != nullptr && g_nccl.world > 1; }

bool dp_fused_available() { return plan_only() ? g_plan_world > 1 && g_plan_fused : g_fused == 1; }

size_t dp_fused_min_elems() {
  static const size_t v = [] {
    const char* e = getenv("CG_DP_FUSED_MIN");
    return e ? (size_t)atoll(e) : (size_t)1 << 18;  // 1 MiB: below that the exchange is latency, not bandwidth
  }();
  return v;
}

bool dp_register_windows(const std::vector<const void*>& ptrs) {
  if (!dp_enabled() || g_nccl.world > DP_MAX_WORLD || g_fused == 0) return false;
  static const bool disabled = [] {
    const char* e = getenv("CG_DP_FUSED");
    return e && atoi(e) == 0;
  }();
  if (disabled) {
    g_fused = 0;
    return false;
  }
  const int world = g_nccl.world, rank = g_nccl.rank;
  if (!g_flags) {
    CG_CUDA(cudaMalloc(&g_flags, DP_FLAG_WORDS * sizeof(uint32_t)));
    CG_CUDA(cudaMemset(g_flags, 0, DP_FLAG_WORDS * sizeof(uint32_t)));
  }
  std::vector<const void*> todo;
  auto want = [&](const void* p) {
    if (!find_ptr(p) && std::find(todo.begin(), todo.end(), p) == todo.end()) todo.push_back(p);
  };
  want(g_flags);
  for (const void* p : ptrs) want(p);

  // this rank's side: unique allocations and (allocation, offset) per buffer
  std::vector<char*> bases;
  std::vector<AllocMsg> allocs;
  std::vector<PtrMsg> pm(todo.size());
  int my_ok = 1;
  for (size_t i = 0; i < todo.size(); i++) {
    char* base = allocation_base(todo[i]);
    if (!base) {
      my_ok = 0;
      base = (char*)todo[i];
    }
    size_t k = std::find(bases.begin(), bases.end(), base) - bases.begin();
    if (k == bases.size()) {
      bases.push_back(base);
      AllocMsg m;
      memset(&m, 0, sizeof m);
      if (my_ok && cudaIpcGetMemHandle(&m.handle, base) == cudaSuccess) m.ok = 1;
      else cudaGetLastError();
      for (const auto& e : g_exported)
        if (e.first == base) m.serial = e.second;
      if (!m.serial) {
        m.serial = g_next_serial++;
        g_exported.push_back({base, m.serial});
      }
      allocs.push_back(m);
    }
    pm[i].alloc = k;
    pm[i].offset = (unsigned long long)((const char*)todo[i] - base);
  }
  // Collective from here on. Ranks build identical plans, so they ask for the same number of buffers; if they ever
  // do not, every rank sees it in the gathered counts and all of them fall back together.
  unsigned long long head[2] = {todo.size(), allocs.size()};
  std::vector<unsigned char> heads = allgather_bytes(head, sizeof head);
  unsigned long long max_allocs = 0;
  for (int r = 0; r < world; r++) {
    unsigned long long h[2];
    memcpy(h, heads.data() + r * sizeof h, sizeof h);
    if (h[0] != head[0]) {
      g_fused = 0;
      return false;
    }
    max_allocs = std::max(max_allocs, h[1]);
  }
  if (todo.empty()) return g_fused == 1;
  const size_t blob_bytes = max_allocs * sizeof(AllocMsg) + todo.size() * sizeof(PtrMsg);
  std::vector<unsigned char> blob(blob_bytes, 0);
  memcpy(blob.data(), allocs.data(), allocs.size() * sizeof(AllocMsg));
  memcpy(blob.data() + max_allocs * sizeof(AllocMsg), pm.data(), pm.size() * sizeof(PtrMsg));
  std::vector<unsigned char> all = allgather_bytes(blob.data(), blob_bytes);

  int good = 1;
  std::vector<PeerMap> maps(todo.size());
  for (size_t i = 0; i < todo.size(); i++) {
    memset(&maps[i], 0, sizeof(PeerMap));
    maps[i].local = todo[i];
    maps[i].peer[rank] = const_cast<void*>(todo[i]);
  }
  std::vector<Opened> fresh;
  for (int r = 0; r < world && good; r++) {
    if (r == rank) continue;
    const unsigned char* rb = all.data() + (size_t)r * blob_bytes;
    unsigned long long n_allocs_r;
    {
      unsigned long long h[2];
      memcpy(h, heads.data() + r * sizeof h, sizeof h);
      n_allocs_r = h[1];
    }
    std::vector<char*> rbase(n_allocs_r, nullptr);
    for (size_t i = 0; i < todo.size() && good; i++) {
      PtrMsg m;
      memcpy(&m, rb + max_allocs * sizeof(AllocMsg) + i * sizeof(PtrMsg), sizeof m);
      if (m.alloc >= n_allocs_r) {
        good = 0;
        break;
      }
      if (!rbase[m.alloc]) {
        AllocMsg am;
        memcpy(&am, rb + m.alloc * sizeof(AllocMsg), sizeof am);
        if (!am.ok) {
          good = 0;
          break;
        }
        for (const Opened& o : g_opened)
          if (o.rank == r && o.serial == am.serial) rbase[m.alloc] = o.base;
        for (const Opened& o : fresh)
          if (o.rank == r && o.serial == am.serial) rbase[m.alloc] = o.base;
        if (!rbase[m.alloc]) {
          // same handle under an older serial: the owner freed that block and the allocator handed the address out
          // again — drop the stale mapping before opening the new one
          for (size_t q = 0; q < g_opened.size();) {
            if (g_opened[q].rank == r && memcmp(&g_opened[q].handle, &am.handle, sizeof am.handle) == 0) {
              cudaIpcCloseMemHandle(g_opened[q].base);
              g_opened.erase(g_opened.begin() + q);
            } else {
              q++;
            }
          }
          void* ptr = nullptr;
          if (cudaIpcOpenMemHandle(&ptr, am.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
            break;
          }
          rbase[m.alloc] = static_cast<char*>(ptr);
          fresh.push_back(Opened{r, am.serial, am.handle, rbase[m.alloc]});
        }
      }
      maps[i].peer[r] = rbase[m.alloc] + m.offset;
    }
  }
  // the decision must be the same everywhere: one more round with this rank's outcome
  unsigned long long outcome = (unsigned long long)good;
  std::vector<unsigned char> outcomes = allgather_bytes(&outcome, sizeof outcome);
  for (int r = 0; r < world; r++) {
    unsigned long long o;
    memcpy(&o, outcomes.data() + r * sizeof o, sizeof o);
    if (!o) good = 0;
  }
  if (!good) {
    for (Opened& o : fresh) cudaIpcCloseMemHandle(o.base);
    g_fused = 0;
    return false;
  }
  for (Opened& o : fresh) g_opened.push_back(o);
  for (PeerMap& m : maps) g_ptrs.push_back(m);
  g_fused = 1;
  return true;
}

void dp_forget(const void* ptr, size_t bytes) {
  if (g_ptrs.empty() && g_exported.empty()) return;
  const char* lo = static_cast<const char*>(ptr);
  for (size_t q = 0; q < g_ptrs.size();) {
    const char* l = static_cast<const char*>(g_ptrs[q].local);
    if (l >= lo && l < lo + (bytes ? bytes : 1)) g_ptrs.erase(g_ptrs.begin() + q);
    else q++;
  }
  char* base = allocation_base(ptr);
  for (size_t q = 0; q < g_exported.size();) {
    if (g_exported[q].first == base) g_exported.erase(g_exported.begin() + q);
    else q++;
  }
}

bool dp_peer_pointers(const void* local, void* peers[DP_MAX_WORLD]) {
  PeerMap* m = find_ptr(local);
  if (!m) return false;
  for (int r = 0; r < DP_MAX_WORLD; r++) peers[r] = r < g_nccl.world ? m->peer[r] : nullptr;
  return true;
}

void dp_fill_flags(DpReduceArgs& a) {
  void* peers[DP_MAX_WORLD];
  if (!g_flags || !dp_peer_pointers(g_flags, peers)) throw Error("internal: data-parallel flag window is not registered");
  for (int r = 0; r < DP_MAX_WORLD; r++) a.flags[r] = static_cast<uint32_t*>(peers[r]);
  a.rank = g_nccl.rank;
  a.world = g_nccl.world;
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("CG_DP_TIMEOUT_MS");  // how long an exchange waits for a peer before it reports a lost rank
    const long long ms = e ? atoll(e) : 0;
    return ms > 0 ? (unsigned long long)ms * 1000000ull : DP_WAIT_TIMEOUT_NS;
  }();
  a.timeout_ns = timeout_ns;
}

void dp_check_errors() {
  if (!g_flags) return;
  uint32_t err = 0;
  CG_CUDA(cudaMemcpy(&err, g_flags + DP_FLAG_ERR, sizeof err, cudaMemcpyDeviceToHost));
  if (err) {
    CG_CUDA(cudaMemset(g_flags + DP_FLAG_ERR, 0, sizeof err));
    throw Error("data-parallel exchange timed out waiting for a peer rank (ranks must run the same iterations in the same order)");
  }
}

void dp_allreduce_sum(float* buf, size_t n, cudaStream_t stream) {
  if (!g_nccl.comm) throw Error("data-parallel all-reduce requested but cg_dp_init was not called");
  nccl_check(g_nccl.AllReduce(buf, buf, n, ncclFloat32, ncclSum, g_nccl.comm, stream), "ncclAllReduce");
}

}  // namespace cg

using namespace cg;
static_assert(CG_DP_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");

extern "C" {
extern const char* cg_last_error(void);
int cg_dp_set_error(const char* msg);  // capi.cpp

int cg_dp_unique_id(unsigned char id[CG_DP_ID_BYTES]) {
  try {
    load_nccl();
    ncclUniqueId uid;
    nccl_check(g_nccl.GetUniqueId(&uid), "ncclGetUniqueId");
    memcpy(id, uid.internal, CG_DP_ID_BYTES);
    return 0;
  } catch (const std::exception& e) {
    return cg_dp_set_error(e.what());
  }
}
int cg_dp_init(int rank, int world, const unsigned char id[CG_DP_ID_BYTES]) {
  try {
    Runtime& rt = Runtime::get();
    load_nccl();
    if (g_nccl.comm) {
      CG_CUDA(cudaStreamSynchronize(rt.stream));
      release_all_plan_graphs();
      close_windows();
      g_nccl.CommDestroy(g_nccl.comm);
      g_nccl.comm = nullptr;
    }
    ncclUniqueId uid;
    memcpy(uid.internal, id, CG_DP_ID_BYTES);
    nccl_check(g_nccl.CommInitRank(&g_nccl.comm, world, uid, rank), "ncclCommInitRank");
    g_nccl.rank = rank;
    g_nccl.world = world;
    rt.dp_epoch++;
    // Decide once, together, whether the ranks can address each other's memory (CUDA IPC over NVLink): opens the
    // flag blocks. Plans built from here on use the fused exchange kernel when they can, ncclAllReduce otherwise.
    if (world > 1) dp_register_windows({});
    // NCCL's all-reduce kernels need SMs of their own to overlap the products of the backward pass, so the persistent
    // product grid leaves some free; the fused exchange kernel is sized to sit BESIDE a product CTA on the same SM
    // (dp_kernels.cu), so nothing is reserved for it.
    int reserve = dp_fused_available() ? 0 : 16;
    if (const char* e = getenv("CG_DP_RESERVE_SMS")) reserve = atoi(e);
    gemm_tf32_set_max_ctas(world > 1 ? kNumSMs - reserve : kNumSMs);
    return 0;
  } catch (const std::exception& e) {
    return cg_dp_set_error(e.what());
  }
}
int cg_dp_fused(void) { return dp_fused_available() ? 1 : 0; }
int cg_dp_plan_only(int world, int fused) {
  try {
    if (!plan_only()) throw Error("cg_dp_plan_only is only available in plan-only mode (cg_set_plan_only)");
    g_plan_world = world;
    g_plan_fused = fused;
    Runtime::get().dp_epoch++;
    return 0;
  } catch (const std::exception& e) {
    return cg_dp_set_error(e.what());
  }
}
int cg_dp_shutdown(void) {
  try {
    if (g_nccl.comm) {
      CG_CUDA(cudaStreamSynchronize(Runtime::get().stream));
      release_all_plan_graphs();  // captured all-reduces pin the communicator: ncclCommDestroy would never return
      close_windows();
      g_nccl.CommDestroy(g_nccl.comm);
      g_nccl.comm = nullptr;
      g_nccl.world = 1;
      gemm_tf32_set_max_ctas(kNumSMs);
      Runtime::get().dp_epoch++;
    }
    return 0;
  } catch (const std::exception& e) {
    return cg_dp_set_error(e.what());
  }
}
}
