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
// From three ranks up the exchange is done INSIDE the NVSwitch instead (NVLS multimem, dp_kernels.cu
// dp_multimem_update_kernel): the gradient slots and the replicated weights of a plan then live in the symmetric
// multicast heap of symm.cpp, which this file initialises next to the communicator.
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
// Freeing an allocation that a peer still has open through cudaIpcOpenMemHandle is undefined behaviour, so a buffer that
// was exported is not freed by its owner's destructor: it waits here until every peer has closed the mapping (the next
// collective point: dp_register_windows or the shutdown).
struct Retired {
  void* ptr;                  // what cudaFree gets
  char* base;                 // the allocation peers have open
  unsigned long long serial;  // its export number
};
std::vector<Retired> g_graveyard;
uint32_t* g_mc_flags = nullptr;     // counter block of the in-switch exchange (in the multicast heap)
uint32_t* g_mc_flags_mc = nullptr;  // the same block through the multicast object
int g_multimem_mode = 1;            // 0 never, 1 from three ranks up (fewer wire bytes than peer memory only there), 2 always
uint32_t* g_flags = nullptr;  // this rank's flag block
int g_fused = -1;             // -1 unknown, 0 unavailable (stay on ncclAllReduce), 1 available

std::vector<unsigned char> allgather_bytes(const void* mine, size_t bytes);

// Collective (every rank calls it at shutdown / re-initialisation).
void close_windows() {
  for (Opened& o : g_opened) cudaIpcCloseMemHandle(o.base);
  g_opened.clear();
  g_ptrs.clear();
  g_exported.clear();
  // nothing exported may be freed before every peer has closed it: one round over the communicator is the barrier
  if (g_nccl.comm && g_nccl.world > 1 && (g_flags || !g_graveyard.empty())) {
    unsigned long long one = 1;
    try {
      allgather_bytes(&one, sizeof one);
    } catch (const std::exception&) {
    }
  }
  for (Retired& r : g_graveyard) cudaFree(r.ptr);
  g_graveyard.clear();
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
  if (!g_nccl.comm) throw Error("internal: data-parallel control message without a communicator");
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

std::vector<unsigned char> dp_allgather_bytes(const void* mine, size_t bytes) { return allgather_bytes(mine, bytes); }
int dp_rank() { return g_nccl.rank; }
int dp_world() { return g_nccl.world; }

// plan-only mode (graph.cpp): a pretended world, so that the data-parallel tape can be inspected without GPUs
static int g_plan_world = 1, g_plan_fused = 0;  // g_plan_fused: 1 peer-memory exchange, 2 in-switch exchange

bool dp_enabled() { return plan_only() ? g_plan_world > 1 : g_nccl.comm != nullptr && g_nccl.world > 1; }

bool dp_fused_available() { return plan_only() ? g_plan_world > 1 && g_plan_fused : (g_fused == 1 || dp_multimem_active()); }

bool dp_multimem_active() {
  if (plan_only()) return g_plan_world > 1 && g_plan_fused == 2;
  if (!symm_available() || !g_mc_flags || g_multimem_mode == 0) return false;
  return g_multimem_mode == 2 || g_nccl.world >= 3;
}

static long long g_fused_min = -1;  // < 0: CG_DP_FUSED_MIN from the environment, else 2^18
size_t dp_fused_min_elems() {
  if (g_fused_min >= 0) return (size_t)g_fused_min;
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
  // Allocations in the graveyard that no live registered buffer shares any more: their export numbers go to the peers,
  // which close their mappings below; the memory is freed once everybody has.
  std::vector<unsigned long long> retire;
  for (const Retired& g : g_graveyard) {
    bool shared = false;
    for (const PeerMap& m : g_ptrs)
      if (allocation_base(m.local) == g.base) shared = true;
    for (const void* t : todo)
      if (allocation_base(t) == g.base) shared = true;
    if (!shared && std::find(retire.begin(), retire.end(), g.serial) == retire.end()) retire.push_back(g.serial);
  }
  // Collective from here on. Ranks build identical plans, so they ask for the same number of buffers; if they ever
  // do not, every rank sees it in the gathered counts and all of them fall back together.
  unsigned long long head[3] = {todo.size(), allocs.size(), retire.size()};
  std::vector<unsigned char> heads = allgather_bytes(head, sizeof head);
  unsigned long long max_allocs = 0, max_retire = 0;
  bool same_count = true;
  for (int r = 0; r < world; r++) {
    unsigned long long h[3];
    memcpy(h, heads.data() + r * sizeof h, sizeof h);
    if (h[0] != head[0]) same_count = false;
    max_allocs = std::max(max_allocs, h[1]);
    max_retire = std::max(max_retire, h[2]);
  }
  if (!same_count) {
    g_fused = 0;
    return false;
  }
  if (todo.empty() && max_retire == 0) return g_fused == 1;
  const size_t retire_at = max_allocs * sizeof(AllocMsg) + todo.size() * sizeof(PtrMsg);
  const size_t blob_bytes = retire_at + max_retire * sizeof(unsigned long long);
  std::vector<unsigned char> blob(blob_bytes ? blob_bytes : 8, 0);
  if (!allocs.empty()) memcpy(blob.data(), allocs.data(), allocs.size() * sizeof(AllocMsg));
  if (!pm.empty()) memcpy(blob.data() + max_allocs * sizeof(AllocMsg), pm.data(), pm.size() * sizeof(PtrMsg));
  if (!retire.empty()) memcpy(blob.data() + retire_at, retire.data(), retire.size() * sizeof(unsigned long long));
  std::vector<unsigned char> all = allgather_bytes(blob.data(), blob.size());
  const size_t blob_stride = blob.size();
  // close what the peers have retired
  for (int r = 0; r < world; r++) {
    if (r == rank) continue;
    unsigned long long h[3];
    memcpy(h, heads.data() + r * sizeof h, sizeof h);
    for (unsigned long long q = 0; q < h[2]; q++) {
      unsigned long long serial;
      memcpy(&serial, all.data() + (size_t)r * blob_stride + retire_at + q * sizeof serial, sizeof serial);
      for (size_t k = 0; k < g_opened.size();) {
        if (g_opened[k].rank == r && g_opened[k].serial == serial) {
          cudaIpcCloseMemHandle(g_opened[k].base);
          g_opened.erase(g_opened.begin() + k);
        } else {
          k++;
        }
      }
    }
  }

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
    const unsigned char* rb = all.data() + (size_t)r * blob_stride;
    unsigned long long n_allocs_r;
    {
      unsigned long long h[3];
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
  // (the outcome round is also the barrier behind which every peer has closed what this rank retired)
  for (size_t q = 0; q < g_graveyard.size();) {
    if (std::find(retire.begin(), retire.end(), g_graveyard[q].serial) != retire.end()) {
      for (size_t e = 0; e < g_exported.size();)
        if (g_exported[e].second == g_graveyard[q].serial) g_exported.erase(g_exported.begin() + e);
        else e++;
      cudaFree(g_graveyard[q].ptr);
      g_graveyard.erase(g_graveyard.begin() + q);
    } else {
      q++;
    }
  }
  if (!good) {
    for (Opened& o : fresh) cudaIpcCloseMemHandle(o.base);
    g_fused = 0;
    return false;
  }
  for (Opened& o : fresh) g_opened.push_back(o);
  for (PeerMap& m : maps) g_ptrs.push_back(m);
  if (!todo.empty() || g_fused == -1) g_fused = 1;
  return g_fused == 1;
}

// A device buffer is about to be freed: drop its registrations. Returns true when this file has taken over the
// cudaFree (the allocation was exported and a peer may still have it open).
bool dp_retire(void* ptr, size_t bytes) {
  if (g_ptrs.empty() && g_exported.empty()) return false;
  const char* lo = static_cast<const char*>(ptr);
  for (size_t q = 0; q < g_ptrs.size();) {
    const char* l = static_cast<const char*>(g_ptrs[q].local);
    if (l >= lo && l < lo + (bytes ? bytes : 1)) g_ptrs.erase(g_ptrs.begin() + q);
    else q++;
  }
  char* base = allocation_base(ptr);
  if (!base || !g_nccl.comm) return false;
  for (const auto& e : g_exported)
    if (e.first == base) {
      g_graveyard.push_back(Retired{ptr, base, e.second});
      return true;
    }
  return false;
}

bool dp_peer_pointers(const void* local, void* peers[DP_MAX_WORLD]) {
  PeerMap* m = find_ptr(local);
  if (!m) return false;
  for (int r = 0; r < DP_MAX_WORLD; r++) peers[r] = r < g_nccl.world ? m->peer[r] : nullptr;
  return true;
}

static unsigned long long dp_timeout_ns();
void dp_fill_flags(DpReduceArgs& a) {
  void* peers[DP_MAX_WORLD];
  if (!g_flags || !dp_peer_pointers(g_flags, peers)) throw Error("internal: data-parallel flag window is not registered");
  for (int r = 0; r < DP_MAX_WORLD; r++) a.flags[r] = static_cast<uint32_t*>(peers[r]);
  a.rank = g_nccl.rank;
  a.world = g_nccl.world;
  a.timeout_ns = dp_timeout_ns();
}

static unsigned long long dp_timeout_ns() {
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("CG_DP_TIMEOUT_MS");  // how long an exchange waits for a peer before it reports a lost rank
    const long long ms = e ? atoll(e) : 0;
    return ms > 0 ? (unsigned long long)ms * 1000000ull : DP_WAIT_TIMEOUT_NS;
  }();
  return timeout_ns;
}

void dp_fill_mc_flags(DpMcArgs& a) {
  if (!g_mc_flags) throw Error("internal: the in-switch exchange has no counter block");
  a.flags_uc = g_mc_flags;
  a.flags_mc = g_mc_flags_mc;
  a.rank = g_nccl.rank;
  a.world = g_nccl.world;
  a.timeout_ns = dp_timeout_ns();
}

void dp_check_errors() {
  for (uint32_t* flags : {g_flags, g_mc_flags}) {
    if (!flags) continue;
    uint32_t err = 0;
    CG_CUDA(cudaMemcpy(&err, flags + DP_FLAG_ERR, sizeof err, cudaMemcpyDeviceToHost));
    if (err) {
      CG_CUDA(cudaMemset(flags + DP_FLAG_ERR, 0, sizeof err));
      throw Error("data-parallel exchange timed out waiting for a peer rank (ranks must run the same iterations in the same order)");
    }
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
      symm_shutdown();
      g_mc_flags = g_mc_flags_mc = nullptr;
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
    // ... and whether the NVSwitch can do the summing (multicast objects over VMM memory, symm.cpp). The job id that
    // names the ranks' unix sockets is a hash of the NCCL unique id: the same on every rank, different per job.
    g_mc_flags = g_mc_flags_mc = nullptr;
    if (world > 1 && world <= DP_MAX_WORLD) {
      unsigned long long job = 1469598103934665603ull;
      for (int k = 0; k < CG_DP_ID_BYTES; k++) job = (job ^ id[k]) * 1099511628211ull;
      int dev = 0;
      CG_CUDA(cudaGetDevice(&dev));
      if (symm_init(job, dev)) {
        g_mc_flags = static_cast<uint32_t*>(symm_alloc(4096));
        void* mc = nullptr;
        symm_locate(g_mc_flags, nullptr, nullptr, &mc);
        g_mc_flags_mc = static_cast<uint32_t*>(mc);
        CG_CUDA(cudaMemset(g_mc_flags, 0, 4096));
        CG_CUDA(cudaDeviceSynchronize());
        unsigned long long one = 1;
        allgather_bytes(&one, sizeof one);  // every rank's counters are zero before anybody's kernel adds to them
      }
    }
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
int cg_dp_set_fused_min(long long min_elems) {
  try {
    if (g_fused_min != min_elems) Runtime::get().dp_epoch++;
    g_fused_min = min_elems;
    return 0;
  } catch (const std::exception& e) {
    return cg_dp_set_error(e.what());
  }
}
int cg_dp_multimem(void) { return dp_multimem_active() ? 1 : 0; }
int cg_dp_multimem_supported(void) { return symm_available() ? 1 : 0; }
const char* cg_dp_multimem_reason(void) { return symm_unavailable_reason(); }
int cg_dp_set_multimem(int mode) {
  try {
    if (mode < 0 || mode > 2) throw Error("cg_dp_set_multimem: mode must be 0 (never), 1 (from three ranks up) or 2 (always)");
    if (g_multimem_mode != mode) Runtime::get().dp_epoch++;  // the exchange kernel is part of a compiled plan
    g_multimem_mode = mode;
    return 0;
  } catch (const std::exception& e) {
    return cg_dp_set_error(e.what());
  }
}
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
      {
        unsigned long long one = 1;
        allgather_bytes(&one, sizeof one);  // nobody unmaps its part of the multicast heap under a peer's last exchange
      }
      symm_shutdown();
      g_mc_flags = g_mc_flags_mc = nullptr;
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
