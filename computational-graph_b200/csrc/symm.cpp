// symm.cpp — the symmetric multicast heap behind the in-switch (NVLS) weight-gradient exchange.
//
// `multimem.ld_reduce` / `multimem.st` (dp_kernels.cu) address a MULTICAST object: one virtual address that names the
// same offset of a physical allocation on every GPU of the NVSwitch domain. A load through it is summed inside the
// switch, a store through it lands on every GPU. That needs (1) physical memory made with the virtual memory
// management API (cuMemCreate, not cudaMalloc), (2) a multicast object every rank has added its device to and bound
// its memory to, and (3) the buffers of one exchange at the SAME offset on every rank.
//
// So the gradient slots and the replicated weights of a data-parallel plan live in this heap: segments of
// (physical memory + multicast object), each mapped twice into this process — at a unicast address (what every
// ordinary kernel uses) and at the multicast address (what the exchange kernel uses) — and carved by a deterministic
// allocator. Ranks run the same program, so they carve the same offsets; bind_dp_updates (plan.cpp) still verifies
// that collectively before a plan may use them. Creating a segment is collective: rank 0 creates the multicast
// object and hands it to the other processes as a POSIX file descriptor over an abstract-namespace unix socket
// (SCM_RIGHTS) — one process per GPU means no shared address space; the control messages around it ride the NCCL
// communicator of dp.cpp (allgather_bytes). Nothing is ever unmapped before cg_dp_shutdown: a freed block goes back to
// a size-keyed free list, so no rank can pull memory from under a peer's exchange.
//
// The reference has no distributed code at all (SURVEY §2.1 last row); this file replaces nothing in it.
#include <cuda.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <map>

#include "graph.h"

namespace cg {

std::vector<unsigned char> dp_allgather_bytes(const void* mine, size_t bytes);  // dp.cpp: over the NCCL communicator
int dp_rank();
int dp_world();

namespace {

// ---- driver entry points (resolved at run time: the library must load on machines without a driver) -------------
struct Driver {
  bool ok = false;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};
Driver g_drv;

template <typename F>
bool resolve(F& fn, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

bool load_driver() {
  static bool tried = false;
  if (tried) return g_drv.ok;
  tried = true;
  Driver& d = g_drv;
  d.ok = resolve(d.DeviceGet, "cuDeviceGet") && resolve(d.DeviceGetAttribute, "cuDeviceGetAttribute") &&
         resolve(d.MulticastCreate, "cuMulticastCreate") && resolve(d.MulticastAddDevice, "cuMulticastAddDevice") &&
         resolve(d.MulticastBindMem, "cuMulticastBindMem") && resolve(d.MulticastUnbind, "cuMulticastUnbind") &&
         resolve(d.MulticastGetGranularity, "cuMulticastGetGranularity") && resolve(d.MemCreate, "cuMemCreate") &&
         resolve(d.MemRelease, "cuMemRelease") && resolve(d.MemGetAllocationGranularity, "cuMemGetAllocationGranularity") &&
         resolve(d.MemAddressReserve, "cuMemAddressReserve") && resolve(d.MemAddressFree, "cuMemAddressFree") &&
         resolve(d.MemMap, "cuMemMap") && resolve(d.MemUnmap, "cuMemUnmap") && resolve(d.MemSetAccess, "cuMemSetAccess") &&
         resolve(d.MemExportToShareableHandle, "cuMemExportToShareableHandle") &&
         resolve(d.MemImportFromShareableHandle, "cuMemImportFromShareableHandle") && resolve(d.GetErrorString, "cuGetErrorString");
  return d.ok;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

// ---- heap state --------------------------------------------------------------------------------------------------
struct Segment {
  size_t bytes = 0, top = 0;
  CUmemGenericAllocationHandle mem = 0, mc = 0;
  CUdeviceptr uc_va = 0, mc_va = 0;
  bool bound = false, uc_mapped = false, mc_mapped = false;
  std::map<size_t, std::vector<size_t>> free_blocks;  // bytes -> offsets (exact-size reuse: weights come in few sizes)
};
struct Block {
  int seg;
  size_t off, bytes;
};

std::vector<Segment> g_segs;
std::map<const void*, Block> g_live;  // unicast pointer -> block
int g_state = -1;                     // -1 not tried, 0 unavailable, 1 available
std::string g_why = "cg_dp_init has not been called with more than one rank";
int g_sock = -1;
unsigned long long g_job = 0;
CUdevice g_dev = 0;
int g_dev_ordinal = 0;
size_t g_gran = 0;  // every segment size and every bind is a multiple of this
constexpr size_t kBlockAlign = 4096;
constexpr size_t kDefaultSegment = (size_t)1 << 30;

// ---- file-descriptor passing between the ranks of one node --------------------------------------------------------
void sock_name(sockaddr_un& a, socklen_t& len, int rank) {
  memset(&a, 0, sizeof a);
  a.sun_family = AF_UNIX;
  // abstract namespace (leading NUL): no file to clean up, gone with the process
  const int n = snprintf(a.sun_path + 1, sizeof a.sun_path - 1, "cg_b200-%016llx-%d", g_job, rank);
  len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}
bool sock_open(int rank) {
  if (g_sock >= 0) close(g_sock);
  g_sock = socket(AF_UNIX, SOCK_DGRAM, 0);
  if (g_sock < 0) return false;
  sockaddr_un a;
  socklen_t len;
  sock_name(a, len, rank);
  if (bind(g_sock, (sockaddr*)&a, len) != 0) {
    close(g_sock);
    g_sock = -1;
    return false;
  }
  return true;
}
bool send_fd(int to_rank, int fd, unsigned long long tag) {
  sockaddr_un a;
  socklen_t len;
  sock_name(a, len, to_rank);
  msghdr msg;
  memset(&msg, 0, sizeof msg);
  iovec io{&tag, sizeof tag};
  char ctl[CMSG_SPACE(sizeof(int))];
  memset(ctl, 0, sizeof ctl);
  msg.msg_name = &a;
  msg.msg_namelen = len;
  msg.msg_iov = &io;
  msg.msg_iovlen = 1;
  msg.msg_control = ctl;
  msg.msg_controllen = sizeof ctl;
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return sendmsg(g_sock, &msg, 0) == (ssize_t)sizeof tag;
}
int recv_fd(unsigned long long want_tag, int timeout_ms) {
  for (;;) {
    pollfd p{g_sock, POLLIN, 0};
    if (poll(&p, 1, timeout_ms) <= 0) return -1;
    unsigned long long tag = 0;
    msghdr msg;
    memset(&msg, 0, sizeof msg);
    iovec io{&tag, sizeof tag};
    char ctl[CMSG_SPACE(sizeof(int))];
    memset(ctl, 0, sizeof ctl);
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    msg.msg_control = ctl;
    msg.msg_controllen = sizeof ctl;
    if (recvmsg(g_sock, &msg, 0) != (ssize_t)sizeof tag) return -1;
    int fd = -1;
    for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c))
      if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) memcpy(&fd, CMSG_DATA(c), sizeof(int));
    if (tag == want_tag) return fd;
    if (fd >= 0) close(fd);  // a stale message of an abandoned attempt
  }
}

// Every rank contributes one word; true when all of them are non-zero. Doubles as the barrier between the phases.
bool all_ok(bool mine) {
  unsigned long long v = mine ? 1 : 0;
  std::vector<unsigned char> all = dp_allgather_bytes(&v, sizeof v);
  for (int r = 0; r < dp_world(); r++) {
    unsigned long long o;
    memcpy(&o, all.data() + (size_t)r * sizeof o, sizeof o);
    if (!o) return false;
  }
  return true;
}

void destroy_segment(Segment& s) {
  Driver& d = g_drv;
  if (s.mc_mapped) d.MemUnmap(s.mc_va, s.bytes);
  if (s.uc_mapped) d.MemUnmap(s.uc_va, s.bytes);
  if (s.mc_va) d.MemAddressFree(s.mc_va, s.bytes);
  if (s.uc_va) d.MemAddressFree(s.uc_va, s.bytes);
  if (s.bound) d.MulticastUnbind(s.mc, g_dev, 0, s.bytes);
  if (s.mem) d.MemRelease(s.mem);
  if (s.mc) d.MemRelease(s.mc);
  s = Segment();
}

// Collective. Adds one segment of at least `min_bytes`; false (on every rank alike) when any rank failed a phase.
bool add_segment(size_t min_bytes, std::string& why) {
  Driver& d = g_drv;
  const int world = dp_world(), rank = dp_rank();
  Segment s;
  s.bytes = (std::max(min_bytes, kDefaultSegment) + g_gran - 1) / g_gran * g_gran;
  const unsigned long long tag = ((unsigned long long)g_segs.size() << 32) | 0x6d63u;
  std::string err;
  auto fail = [&](const std::string& m) {
    if (err.empty()) err = m;
    return false;
  };
  // 1. rank 0 creates the multicast object and passes it on; everyone ends up holding a handle to it
  bool ok = true;
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof mp);
  mp.numDevices = (unsigned)world;
  mp.size = s.bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  if (rank == 0) {
    CUresult r = d.MulticastCreate(&s.mc, &mp);
    if (r != CUDA_SUCCESS) ok = fail("cuMulticastCreate: " + cu_err(r));
    int fd = -1;
    if (ok) {
      r = d.MemExportToShareableHandle(&fd, s.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
      if (r != CUDA_SUCCESS) ok = fail("cuMemExportToShareableHandle(multicast): " + cu_err(r));
    }
    if (ok)
      for (int p = 1; p < world; p++)
        if (!send_fd(p, fd, tag)) ok = fail("sendmsg(SCM_RIGHTS) to a peer rank failed");
    if (fd >= 0) close(fd);
  }
  if (!all_ok(ok)) {  // (also: rank 0 has sent before anybody blocks in recv_fd)
    destroy_segment(s);
    why = err.empty() ? "a peer rank could not create or export the multicast object" : err;
    return false;
  }
  if (rank != 0) {
    const int fd = recv_fd(tag, 30000);
    if (fd < 0) ok = fail("did not receive the multicast object's file descriptor");
    if (ok) {
      CUresult r = d.MemImportFromShareableHandle(&s.mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      if (r != CUDA_SUCCESS) ok = fail("cuMemImportFromShareableHandle(multicast): " + cu_err(r));
    }
    if (fd >= 0) close(fd);
  }
  // 2. every device joins the team (binding blocks until all have)
  if (ok) {
    CUresult r = d.MulticastAddDevice(s.mc, g_dev);
    if (r != CUDA_SUCCESS) ok = fail("cuMulticastAddDevice: " + cu_err(r));
  }
  if (!all_ok(ok)) {
    destroy_segment(s);
    why = err.empty() ? "a peer rank could not join the multicast team" : err;
    return false;
  }
  // 3. this rank's physical memory, bound at offset 0, mapped at a unicast and at the multicast address
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = g_dev_ordinal;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;  // shareable multicast objects bind only shareable memory
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUresult r = d.MemCreate(&s.mem, s.bytes, &prop, 0);
  if (r != CUDA_SUCCESS) ok = fail("cuMemCreate: " + cu_err(r));
  if (ok && (r = d.MulticastBindMem(s.mc, 0, s.mem, 0, s.bytes, 0)) != CUDA_SUCCESS) ok = fail("cuMulticastBindMem: " + cu_err(r));
  if (ok) s.bound = true;
  if (ok && (r = d.MemAddressReserve(&s.uc_va, s.bytes, g_gran, 0, 0)) != CUDA_SUCCESS) ok = fail("cuMemAddressReserve: " + cu_err(r));
  if (ok && (r = d.MemMap(s.uc_va, s.bytes, 0, s.mem, 0)) != CUDA_SUCCESS) ok = fail("cuMemMap(unicast): " + cu_err(r));
  if (ok) s.uc_mapped = true;
  if (ok && (r = d.MemSetAccess(s.uc_va, s.bytes, &acc, 1)) != CUDA_SUCCESS) ok = fail("cuMemSetAccess(unicast): " + cu_err(r));
  if (ok && (r = d.MemAddressReserve(&s.mc_va, s.bytes, g_gran, 0, 0)) != CUDA_SUCCESS) ok = fail("cuMemAddressReserve: " + cu_err(r));
  if (ok && (r = d.MemMap(s.mc_va, s.bytes, 0, s.mc, 0)) != CUDA_SUCCESS) ok = fail("cuMemMap(multicast): " + cu_err(r));
  if (ok) s.mc_mapped = true;
  if (ok && (r = d.MemSetAccess(s.mc_va, s.bytes, &acc, 1)) != CUDA_SUCCESS) ok = fail("cuMemSetAccess(multicast): " + cu_err(r));
  if (!all_ok(ok)) {  // nobody may touch the multicast address before every rank has bound its memory
    destroy_segment(s);
    why = err.empty() ? "a peer rank could not bind or map its memory" : err;
    return false;
  }
  g_segs.push_back(std::move(s));
  return true;
}

}  // namespace

bool symm_available() { return g_state == 1; }
const char* symm_unavailable_reason() { return g_why.c_str(); }

// Collective (cg_dp_init): decides, on every rank alike, whether the multicast heap can be used, and creates its first
// segment — whose first block holds the barrier counters of the in-switch exchange kernel.
bool symm_init(unsigned long long job_id, int device_ordinal) {
  symm_shutdown();
  g_job = job_id;
  g_dev_ordinal = device_ordinal;
  bool ok = true;
  std::string why;
  if (const char* e = getenv("CG_DP_MULTIMEM"))
    if (atoi(e) == 0) ok = false, why = "disabled by CG_DP_MULTIMEM=0";
  if (ok && !load_driver()) ok = false, why = "the CUDA driver does not export the multicast API";
  if (ok) {
    int mc = 0, vmm = 0;
    Driver& d = g_drv;
    if (d.DeviceGet(&g_dev, device_ordinal) != CUDA_SUCCESS) ok = false, why = "cuDeviceGet failed";
    if (ok) d.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, g_dev);
    if (ok) d.DeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, g_dev);
    if (ok && !mc) ok = false, why = "the device reports no multicast support (no NVSwitch fabric)";
    if (ok && !vmm) ok = false, why = "the device cannot export memory as POSIX file descriptors";
    if (ok) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof mp);
      mp.numDevices = (unsigned)dp_world();
      mp.size = kDefaultSegment;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t g_min = 0, g_rec = 0, g_mem = 0;
      CUmemAllocationProp prop;
      memset(&prop, 0, sizeof prop);
      prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id = device_ordinal;
      prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      if (d.MulticastGetGranularity(&g_min, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS ||
          d.MulticastGetGranularity(&g_rec, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS ||
          d.MemGetAllocationGranularity(&g_mem, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS)
        ok = false, why = "could not query the multicast granularity";
      else
        g_gran = std::max(std::max(g_min, g_rec), std::max(g_mem, (size_t)1 << 21));
    }
  }
  if (ok && !sock_open(dp_rank())) ok = false, why = "could not bind the unix socket used to pass the multicast handle";
  // everybody has a bound socket (or the whole group gives up) before rank 0 sends anything
  if (!all_ok(ok)) {
    g_state = 0;
    g_why = why.empty() ? "a peer rank cannot use multicast memory" : why;
    if (g_sock >= 0) close(g_sock), g_sock = -1;
    return false;
  }
  // the granularity must agree, or the segment sizes (and with them every offset) would not
  {
    unsigned long long g = g_gran, gmax = 0;
    std::vector<unsigned char> all = dp_allgather_bytes(&g, sizeof g);
    for (int r = 0; r < dp_world(); r++) {
      unsigned long long o;
      memcpy(&o, all.data() + (size_t)r * sizeof o, sizeof o);
      gmax = std::max(gmax, o);
    }
    g_gran = (size_t)gmax;
  }
  if (!add_segment(kDefaultSegment, why)) {
    g_state = 0;
    g_why = why;
    close(g_sock), g_sock = -1;
    return false;
  }
  g_state = 1;
  g_why.clear();
  return true;
}

void symm_shutdown() {
  for (Segment& s : g_segs) destroy_segment(s);
  g_segs.clear();
  g_live.clear();
  if (g_sock >= 0) close(g_sock);
  g_sock = -1;
  g_state = -1;
  g_why = "cg_dp_init has not been called with more than one rank";
}

// Every rank calls this with the same sizes in the same order (a new segment is a collective). Returns the unicast
// address. Throws when the heap cannot grow — on every rank alike.
void* symm_alloc(size_t bytes) {
  if (g_state != 1) throw Error("internal: symmetric heap is not available");
  bytes = (std::max<size_t>(bytes, 1) + kBlockAlign - 1) / kBlockAlign * kBlockAlign;
  for (int pass = 0; pass < 2; pass++) {
    for (size_t k = 0; k < g_segs.size(); k++) {
      Segment& s = g_segs[k];
      auto it = s.free_blocks.find(bytes);
      size_t off;
      if (it != s.free_blocks.end() && !it->second.empty()) {
        off = it->second.back();
        it->second.pop_back();
      } else if (s.top + bytes <= s.bytes) {
        off = s.top;
        s.top += bytes;
      } else {
        continue;
      }
      void* p = reinterpret_cast<void*>(s.uc_va + off);
      g_live[p] = Block{(int)k, off, bytes};
      return p;
    }
    std::string why;
    if (pass == 1 || !add_segment(bytes, why))
      throw Error("data parallel: the multicast heap cannot grow by " + std::to_string(bytes >> 20) + " MiB (" + why +
                  "); CG_DP_MULTIMEM=0 selects the peer-memory exchange");
  }
  return nullptr;
}

void symm_free(void* uc) {
  auto it = g_live.find(uc);
  if (it == g_live.end()) return;  // the heap was shut down under it: the mapping is gone with the segment
  g_segs[it->second.seg].free_blocks[it->second.bytes].push_back(it->second.off);
  g_live.erase(it);
}

// (segment, offset) and the multicast address of a pointer INSIDE a live block; false when it is not heap memory.
bool symm_locate(const void* uc, unsigned long long* seg, unsigned long long* off, void** mc) {
  if (g_live.empty()) return false;
  auto it = g_live.upper_bound(uc);
  if (it == g_live.begin()) return false;
  --it;
  const char* base = static_cast<const char*>(it->first);
  const char* p = static_cast<const char*>(uc);
  if (p < base || p >= base + it->second.bytes) return false;
  const Segment& s = g_segs[it->second.seg];
  const size_t o = it->second.off + (size_t)(p - base);
  if (seg) *seg = (unsigned long long)it->second.seg;
  if (off) *off = o;
  if (mc) *mc = reinterpret_cast<void*>(s.mc_va + o);
  return true;
}

size_t symm_bytes_mapped() {
  size_t t = 0;
  for (const Segment& s : g_segs) t += s.bytes;
  return t;
}

}  // namespace cg
