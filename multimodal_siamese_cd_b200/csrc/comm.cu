// Library-owned NCCL communicator (include/b200cd.h: b200cd_comm_*, b200cd_allreduce_*): the gradient buckets and the
// loss partial sums of the data-parallel step are all-reduced (SUM) over NVLink / NVSwitch without going through
// torch.distributed. NCCL is resolved at run time with dlopen — the library has no link-time dependency on it — from
// the copy already loaded in the process (torch's) or from the path given to b200cd_comm_load.
//
// Replaces (reference): the gather / reduce-add of nn.DataParallel (utils/networks.py:27).
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/b200cd.h"
#include "kernels.h"

namespace {

struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*CommDestroyFn)(NcclComm);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);

constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn get_error_string = nullptr;
  GetVersionFn get_version = nullptr;
} g_nccl;

NcclComm g_comm = nullptr;
int g_rank = -1, g_nranks = 0;
thread_local std::string g_comm_error;

int comm_fail(int code, const std::string& msg) {
  g_comm_error = msg;
  b200cd::set_last_error(msg.c_str());
  return code;
}

int load_api(const char* path) {
  if (g_nccl.handle != nullptr) return 0;
  const char* candidates[] = {path, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* c : candidates) {
    if (c == nullptr || c[0] == 0) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h != nullptr) break;
  }
  if (h == nullptr) return comm_fail(B200CD_ERR_CUDA, std::string("NCCL could not be loaded: ") + (dlerror() ? dlerror() : "?"));
  g_nccl.handle = h;
  g_nccl.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
  g_nccl.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
  g_nccl.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
  g_nccl.get_error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
  g_nccl.get_version = reinterpret_cast<GetVersionFn>(dlsym(h, "ncclGetVersion"));
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy) {
    g_nccl = NcclApi();
    return comm_fail(B200CD_ERR_CUDA, "the NCCL library found lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce");
  }
  return 0;
}

int nccl_check(int rc, const char* what) {
  if (rc == 0) return 0;
  const char* s = g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?";
  return comm_fail(B200CD_ERR_CUDA, std::string(what) + ": NCCL error " + std::to_string(rc) + " (" + s + ")");
}

}  // namespace

extern "C" {

int b200cd_comm_load(const char* libnccl_path) { return load_api(libnccl_path); }

int b200cd_comm_version(void) {
  if (load_api(nullptr) != 0 || g_nccl.get_version == nullptr) return -1;
  int v = 0;
  return g_nccl.get_version(&v) == 0 ? v : -1;
}

int b200cd_comm_unique_id(void* id128) {
  if (id128 == nullptr) return comm_fail(B200CD_ERR_SHAPE, "comm_unique_id: NULL buffer");
  if (int rc = load_api(nullptr)) return rc;
  NcclUniqueId id;
  if (int rc = nccl_check(g_nccl.get_unique_id(&id), "ncclGetUniqueId")) return rc;
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int b200cd_comm_init(const void* id128, int rank, int nranks) {
  if (id128 == nullptr || nranks < 1 || rank < 0 || rank >= nranks) return comm_fail(B200CD_ERR_SHAPE, "comm_init: bad rank / nranks");
  if (g_comm != nullptr) return comm_fail(B200CD_ERR_SHAPE, "comm_init: a communicator already exists (b200cd_comm_destroy first)");
  if (int rc = load_api(nullptr)) return rc;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  if (int rc = nccl_check(g_nccl.comm_init_rank(&g_comm, nranks, id, rank), "ncclCommInitRank")) {
    g_comm = nullptr;
    return rc;
  }
  g_rank = rank;
  g_nranks = nranks;
  return 0;
}

int b200cd_comm_size(void) { return g_comm != nullptr ? g_nranks : 0; }

int b200cd_allreduce_bucket(float* buf, int64_t count, void* stream) {
  if (g_comm == nullptr) return comm_fail(B200CD_ERR_SHAPE, "allreduce_bucket: b200cd_comm_init has not been called");
  if (count == 0) return 0;  // an empty bucket (a backward segment that completes no gradient) is a no-op on every rank
  if (buf == nullptr || count < 0) return comm_fail(B200CD_ERR_SHAPE, "allreduce_bucket: NULL buffer / negative count");
  return nccl_check(g_nccl.all_reduce(buf, buf, static_cast<size_t>(count), kNcclFloat32, kNcclSum, g_comm,
                                      reinterpret_cast<cudaStream_t>(stream)), "ncclAllReduce(f32)");
}

int b200cd_allreduce_f64(double* buf, int64_t count, void* stream) {
  if (g_comm == nullptr) return comm_fail(B200CD_ERR_SHAPE, "allreduce_f64: b200cd_comm_init has not been called");
  if (count == 0) return 0;
  if (buf == nullptr || count < 0) return comm_fail(B200CD_ERR_SHAPE, "allreduce_f64: NULL buffer / negative count");
  return nccl_check(g_nccl.all_reduce(buf, buf, static_cast<size_t>(count), kNcclFloat64, kNcclSum, g_comm,
                                      reinterpret_cast<cudaStream_t>(stream)), "ncclAllReduce(f64)");
}

int b200cd_comm_destroy(void) {
  if (g_comm == nullptr) return 0;
  const int rc = nccl_check(g_nccl.comm_destroy(g_comm), "ncclCommDestroy");
  g_comm = nullptr;
  g_rank = -1;
  g_nranks = 0;
  return rc;
}

}  // extern "C"
