// The one collective of the sampling path: an all-gather of the decoded images over NCCL
// (NVLink 5 / NVSwitch), issued on the handle's own stream.  There is no collective inside the DDIM loop.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch bundles, or the system one), so that
// libldm_b200.so itself keeps no dependency on it -- single-GPU users never load NCCL.  Only the five
// entry points below are used; their prototypes are restated from nccl.h (stable ABI since NCCL 2.0).
#include "../../include/ldm_b200.h"
#include "model.h"
#include <dlfcn.h>
#include <mutex>

namespace ldm {

namespace {
typedef struct { char internal[128]; } nccl_unique_id;   // ncclUniqueId, NCCL_UNIQUE_ID_BYTES = 128
typedef void* nccl_comm_t;
enum { NCCL_FLOAT32 = 7 };                                // ncclFloat32

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string path;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

NcclApi& nccl(const char* hint) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return g_nccl;
  std::vector<std::string> cands;
  if (hint && hint[0]) cands.push_back(hint);
  if (const char* e = getenv("LDM_B200_NCCL_LIB")) cands.push_back(e);
  cands.push_back("libnccl.so.2");
  cands.push_back("libnccl.so");
  std::string tried;
  for (const std::string& c : cands) {
    void* l = dlopen(c.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (l) { g_nccl.lib = l; g_nccl.path = c; break; }
    const char* why = dlerror();   // one call: dlerror() clears the message it returns
    tried += c + " (" + (why ? why : "?") + "); ";
  }
  LDM_CHECK(g_nccl.lib, "NCCL not found: tried %s", tried.c_str());
  auto sym = [&](const char* name) {
    void* p = dlsym(g_nccl.lib, name);
    LDM_CHECK(p, "%s: symbol %s missing", g_nccl.path.c_str(), name);
    return p;
  };
  g_nccl.GetUniqueId = reinterpret_cast<int (*)(nccl_unique_id*)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<int (*)(nccl_comm_t*, int, nccl_unique_id, int)>(sym("ncclCommInitRank"));
  g_nccl.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t)>(sym("ncclAllGather"));
  g_nccl.CommDestroy = reinterpret_cast<int (*)(nccl_comm_t)>(sym("ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  g_nccl.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
  return g_nccl;
}

#define NCCL_CHECK(api, expr)                                                                        \
  do {                                                                                               \
    const int _r = (expr);                                                                           \
    if (_r != 0) throw ::ldm::Error(::ldm::fmt("%s:%d: NCCL error %d: %s", __FILE__, __LINE__, _r,   \
                                               (api).GetErrorString ? (api).GetErrorString(_r) : "?")); \
  } while (0)
}  // namespace

void comm_unique_id(const char* lib_hint, char out[128]) {
  NcclApi& a = nccl(lib_hint);
  nccl_unique_id id;
  NCCL_CHECK(a, a.GetUniqueId(&id));
  memcpy(out, id.internal, 128);
}

void Model::comm_init(const char* lib_hint, const char id_bytes[128], int rank, int world) {
  LDM_CHECK(world >= 1 && rank >= 0 && rank < world, "comm_init: rank %d of %d", rank, world);
  comm_destroy();
  NcclApi& a = nccl(lib_hint);
  CUDA_CHECK(cudaSetDevice(eng.device));
  nccl_unique_id id;
  memcpy(id.internal, id_bytes, 128);
  nccl_comm_t c = nullptr;
  NCCL_CHECK(a, a.CommInitRank(&c, world, id, rank));
  comm_ = c; comm_rank_ = rank; comm_world_ = world;
}

void Model::comm_destroy() {
  if (!comm_) return;
  eng.sync();
  g_nccl.CommDestroy(comm_);
  comm_ = nullptr; comm_world_ = 1; comm_rank_ = 0;
}

// Every rank contributes `count` floats (its shard, padded by the caller to the largest shard) and receives
// world * count floats in rank order.  Host pointers are staged through the handle's grow-only buffers.
void Model::allgather(const float* local, long long count, float* global) {
  LDM_CHECK(comm_ != nullptr, "allgather: no communicator (ldm_comm_init)");
  CUDA_CHECK(cudaSetDevice(eng.device));
  auto on_device = [&](const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
  };
  const size_t bytes = (size_t)count * sizeof(float);
  const float* send = local;
  float* recv = global;
  if (!on_device(local)) {
    float* s = static_cast<float*>(stage(ST_A, bytes));
    CUDA_CHECK(cudaMemcpyAsync(s, local, bytes, cudaMemcpyDefault, eng.stream));
    send = s;
  }
  const bool recv_dev = on_device(global);
  if (!recv_dev) recv = static_cast<float*>(stage(ST_B, bytes * comm_world_));
  ensure_events();
  CUDA_CHECK(cudaEventRecord(ev0_, eng.stream));
  NCCL_CHECK(g_nccl, g_nccl.AllGather(send, recv, (size_t)count, NCCL_FLOAT32, comm_, eng.stream));
  CUDA_CHECK(cudaEventRecord(ev1_, eng.stream));
  if (!recv_dev) CUDA_CHECK(cudaMemcpyAsync(global, recv, bytes * comm_world_, cudaMemcpyDefault, eng.stream));
  eng.sync();
  CUDA_CHECK(cudaEventElapsedTime(&last_gather_ms, ev0_, ev1_));
  eng.launches++;   // NCCL's all-gather kernel (library code, not ours: not counted as a GEMM / hot kernel)
}

}  // namespace ldm

using namespace ldm;

#define COMM_API_BEGIN try {
#define COMM_API_END                                     \
  return LDM_OK;                                         \
  }                                                      \
  catch (const std::exception& e) {                      \
    ldm_set_error(e.what());                             \
    cudaGetLastError();                                  \
    return LDM_ERR_CUDA;                                 \
  }

extern "C" {

int ldm_comm_unique_id(const char* nccl_lib, char id_out[128]) {
  COMM_API_BEGIN
  LDM_CHECK(id_out, "ldm_comm_unique_id: null argument");
  comm_unique_id(nccl_lib, id_out);
  COMM_API_END
}

int ldm_comm_init(ldm_handle* h, const char* nccl_lib, const char id[128], int rank, int world) {
  COMM_API_BEGIN
  LDM_CHECK(h && ldm_handle_model(h) && id, "ldm_comm_init: null argument");
  ldm_handle_model(h)->comm_init(nccl_lib, id, rank, world);
  COMM_API_END
}

int ldm_allgather_images(ldm_handle* h, const float* local, int64_t count_per_rank, float* global) {
  COMM_API_BEGIN
  LDM_CHECK(h && ldm_handle_model(h) && local && global && count_per_rank > 0, "ldm_allgather_images: bad argument");
  ldm_handle_model(h)->allgather(local, count_per_rank, global);
  COMM_API_END
}

int ldm_comm_destroy(ldm_handle* h) {
  COMM_API_BEGIN
  LDM_CHECK(h && ldm_handle_model(h), "ldm_comm_destroy: null handle");
  ldm_handle_model(h)->comm_destroy();
  COMM_API_END
}

}  // extern "C"
