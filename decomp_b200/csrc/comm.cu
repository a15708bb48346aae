// Thin NCCL wrappers of the C ABI (comm_init / allreduce / destroy), so that a host without
// torch.distributed can run the sharded solves: the collectives of the hot path -- all-reduce of the [k,f] / [k,k]
// statistics (NMF, dictionary learning), MIN-all-reduce of the convergence latch (Lasso), reduce-scatter of the masked
// [k,f,k] statistic along f and all-gather of the new dictionary (dictionary learning) -- on the caller's stream.
// NCCL is not linked: the library the process already uses (libnccl.so.2, the one torch loads) is bound at run time,
// so both see the same NCCL.  The 128-byte unique id is created on one rank and distributed by the host.
#include <dlfcn.h>

#include <mutex>

#include "common.h"

namespace dcp {

struct NcclId {
  char internal[128];
};
typedef void* NcclComm;
enum { kNcclSum = 0, kNcclMin = 3, kNcclInt32 = 2, kNcclFloat64 = 8 };

struct NcclApi {
  int (*GetUniqueId)(NcclId*);
  int (*CommInitRank)(NcclComm*, int, NcclId, int);
  int (*CommDestroy)(NcclComm);
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
  int (*ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
  const char* (*GetErrorString)(int);
  bool ok;
};

static NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    api.ok = false;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already uses, if any
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW);
    if (h == nullptr) return;
    api.GetUniqueId = reinterpret_cast<int (*)(NcclId*)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclId, int)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)>(
        dlsym(h, "ncclAllReduce"));
    api.ReduceScatter = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)>(
        dlsym(h, "ncclReduceScatter"));
    api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(
        dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.ReduceScatter &&
             api.AllGather && api.GetErrorString;
  });
  return &api;
}

static int check_nccl(int rc, const char* what) {
  if (rc == 0) return DECOMP_OK;
  set_error("%s: NCCL error %d (%s)", what, rc, nccl()->GetErrorString ? nccl()->GetErrorString(rc) : "?");
  return DECOMP_ERR_CUDA;
}

static int need_nccl() {
  if (nccl()->ok) return DECOMP_OK;
  set_error("libnccl.so.2 could not be loaded (decomp_comm_* need NCCL at run time)");
  return DECOMP_ERR_UNSUPPORTED;
}

}  // namespace dcp

using namespace dcp;

extern "C" {

int decomp_comm_unique_id(void* id_out) {
  if (id_out == nullptr) return DECOMP_ERR_INVALID;
  int rc = need_nccl();
  if (rc != DECOMP_OK) return rc;
  return check_nccl(nccl()->GetUniqueId(reinterpret_cast<NcclId*>(id_out)), "decomp_comm_unique_id");
}

int decomp_comm_init(const void* id, int32_t nranks, int32_t rank, void** comm_out) {
  if (id == nullptr || comm_out == nullptr || nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("decomp_comm_init: invalid argument");
    return DECOMP_ERR_INVALID;
  }
  int rc = need_nccl();
  if (rc != DECOMP_OK) return rc;
  NcclId copy = *reinterpret_cast<const NcclId*>(id);
  NcclComm comm = nullptr;
  rc = check_nccl(nccl()->CommInitRank(&comm, nranks, copy, rank), "decomp_comm_init");
  *comm_out = comm;
  return rc;
}

int decomp_comm_destroy(void* comm) {
  if (comm == nullptr) return DECOMP_OK;
  int rc = need_nccl();
  if (rc != DECOMP_OK) return rc;
  return check_nccl(nccl()->CommDestroy(comm), "decomp_comm_destroy");
}

int decomp_comm_allreduce_sum_f64(void* comm, double* buf, int64_t count, void* stream) {
  if (count <= 0) return DECOMP_OK;
  int rc = need_nccl();
  if (rc != DECOMP_OK) return rc;
  return check_nccl(nccl()->AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, comm, as_stream(stream)),
                    "decomp_comm_allreduce_sum_f64");
}

int decomp_comm_allreduce_min_i32(void* comm, int32_t* buf, int64_t count, void* stream) {
  if (count <= 0) return DECOMP_OK;
  int rc = need_nccl();
  if (rc != DECOMP_OK) return rc;
  return check_nccl(nccl()->AllReduce(buf, buf, (size_t)count, kNcclInt32, kNcclMin, comm, as_stream(stream)),
                    "decomp_comm_allreduce_min_i32");
}

int decomp_comm_reduce_scatter_sum_f64(void* comm, const double* send, double* recv, int64_t recv_count, void* stream) {
  if (recv_count <= 0) return DECOMP_OK;
  int rc = need_nccl();
  if (rc != DECOMP_OK) return rc;
  return check_nccl(
      nccl()->ReduceScatter(send, recv, (size_t)recv_count, kNcclFloat64, kNcclSum, comm, as_stream(stream)),
      "decomp_comm_reduce_scatter_sum_f64");
}

int decomp_comm_allgather_f64(void* comm, const double* send, double* recv, int64_t send_count, void* stream) {
  if (send_count <= 0) return DECOMP_OK;
  int rc = need_nccl();
  if (rc != DECOMP_OK) return rc;
  return check_nccl(nccl()->AllGather(send, recv, (size_t)send_count, kNcclFloat64, comm, as_stream(stream)),
                    "decomp_comm_allgather_f64");
}

}  // extern "C"
