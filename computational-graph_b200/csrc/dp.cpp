// dp.cpp — data-parallel plumbing: one process per GPU, one NCCL communicator over the
// NVLink 5 / NVSwitch domain, a SUM all-reduce of every weight gradient before its update (the
// reference has no batch normalisation of the loss — root error == 1, optimize.rs:69 — so gradients
// add, they are not averaged). The reference itself has no distributed code at all.
//
// NCCL is resolved with dlopen at cg_dp_init time, so the library loads (and every single-GPU path
// runs) on machines without NCCL, and a process that already carries torch's NCCL reuses that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "../../include/cg_graph.h"
#include "tape.h"

namespace cg {

namespace {
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
  NCCL_SYM(CommDestroy, "ncclCommDestroy")
  NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
}
void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw Error(std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
}
}  // namespace

bool dp_enabled() { return g_nccl.comm != nullptr && g_nccl.world > 1; }

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
      g_nccl.CommDestroy(g_nccl.comm);
      g_nccl.comm = nullptr;
    }
    ncclUniqueId uid;
    memcpy(uid.internal, id, CG_DP_ID_BYTES);
    nccl_check(g_nccl.CommInitRank(&g_nccl.comm, world, uid, rank), "ncclCommInitRank");
    g_nccl.rank = rank;
    g_nccl.world = world;
    // leave some SMs to the all-reduce kernels so that they overlap the products of the backward pass
    int reserve = 16;
    if (const char* e = getenv("CG_DP_RESERVE_SMS")) reserve = atoi(e);
    gemm_tf32_set_max_ctas(world > 1 ? kNumSMs - reserve : kNumSMs);
    rt.dp_epoch++;
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
