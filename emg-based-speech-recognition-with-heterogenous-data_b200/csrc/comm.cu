// sst_comm_*: the gradient exchange of the data-parallel path behind the C ABI -- replaces the replica
// gather of nn.DataParallel (recognition_model.py:284).  One NCCL communicator per process, created from a unique id the caller
// distributes (any host channel: torch.distributed's store, a file, MPI); sst_comm_allreduce_bucket only ENQUEUES on the given
// stream (NVLink 5 / NVSwitch, NVLS in-switch reduction where NCCL enables it).  NCCL is loaded at run time with dlopen -- the
// same libnccl.so.2 the host framework uses can be named (SST_NCCL_LIB or the `lib_path` argument) -- so libsst.so has no
// link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <stdlib.h>

#include "sst_common.cuh"

namespace sst {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUM = 0, NCCL_AVG = 4 };                 // ncclRedOp_t (nccl.h, stable since 2.10)
enum { NCCL_FLOAT32 = 7, NCCL_BFLOAT16 = 9 };        // ncclDataType_t

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl(const char* lib_path) {
  if (g_nccl.handle != nullptr) return SST_OK;
  const char* env = getenv("SST_NCCL_LIB");
  const char* names[] = {lib_path, env, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (n == nullptr || n[0] == 0) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h != nullptr) break;
  }
  SST_REQUIRE(h != nullptr, SST_E_COMM, "sst_comm: cannot load NCCL (%s); name it with SST_NCCL_LIB", dlerror());
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
  SST_REQUIRE(a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString, SST_E_COMM,
              "sst_comm: the loaded library does not export the NCCL entry points");
  g_nccl = a;
  return SST_OK;
}

struct Comm { ncclComm_t comm; int rank, world; };

}  // namespace sst

using namespace sst;

#define SST_NCCL(call, what)                                                                            \
  do { int r_ = (call); SST_REQUIRE(r_ == 0, SST_E_COMM, "%s: %s", what, g_nccl.GetErrorString(r_)); } while (0)

extern "C" {

int sst_comm_unique_id(void* id_out, const char* lib_path) {
  SST_REQUIRE(id_out != nullptr, SST_E_ARG, "sst_comm_unique_id: null output");
  int rc = load_nccl(lib_path);
  if (rc) return rc;
  ncclUniqueId id;
  SST_NCCL(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
  return SST_OK;
}

int sst_comm_init(const void* id, int rank, int world, const char* lib_path, void** comm_out) {
  SST_REQUIRE(id != nullptr && comm_out != nullptr && world >= 1 && rank >= 0 && rank < world, SST_E_ARG, "sst_comm_init: bad arguments");
  int rc = load_nccl(lib_path);
  if (rc) return rc;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  Comm* c = new Comm{nullptr, rank, world};
  int r = g_nccl.CommInitRank(&c->comm, world, uid, rank);      // uses the calling thread's current CUDA device
  if (r != 0) {
    delete c;
    SST_REQUIRE(false, SST_E_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  }
  *comm_out = c;
  return SST_OK;
}

int sst_comm_allreduce_bucket(void* comm, void* buf, int64_t count, int dtype, int average, void* stream) {
  SST_REQUIRE(comm != nullptr && buf != nullptr && count >= 0, SST_E_ARG, "sst_comm_allreduce_bucket: bad arguments");
  SST_REQUIRE(dtype == SST_F32 || dtype == SST_BF16, SST_E_ARG, "sst_comm_allreduce_bucket: dtype must be SST_F32 or SST_BF16");
  if (count == 0) return SST_OK;
  Comm* c = reinterpret_cast<Comm*>(comm);
  SST_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, dtype == SST_F32 ? NCCL_FLOAT32 : NCCL_BFLOAT16, average ? NCCL_AVG : NCCL_SUM,
                            c->comm, reinterpret_cast<cudaStream_t>(stream)), "ncclAllReduce");
  return SST_OK;
}

int sst_comm_destroy(void* comm) {
  if (comm == nullptr) return SST_OK;
  Comm* c = reinterpret_cast<Comm*>(comm);
  int r = g_nccl.handle != nullptr ? g_nccl.CommDestroy(c->comm) : 0;
  delete c;
  SST_REQUIRE(r == 0, SST_E_COMM, "ncclCommDestroy: %s", g_nccl.GetErrorString(r));
  return SST_OK;
}

int sst_comm_nccl_version(const char* lib_path) {       // e.g. 22809; negative SST_E_* when NCCL cannot be loaded
  int rc = load_nccl(lib_path);
  if (rc) return rc;
  int v = 0;
  if (g_nccl.GetVersion == nullptr || g_nccl.GetVersion(&v) != 0) return 0;
  return v;
}

}  // extern "C"
