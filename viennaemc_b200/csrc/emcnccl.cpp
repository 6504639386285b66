// libemcnccl.so: ncclAllReduce behind the all-reduce callback of the sharded device run (include/emcnccl.h).
#include "../../include/emcnccl.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>

struct emcnccl_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  int64_t calls = 0, bytes = 0;
  double *staging = nullptr;
  int64_t stagingCount = 0;
};

namespace {
thread_local std::string g_error;
int failNccl(const char *what, ncclResult_t r) {
  g_error = std::string(what) + ": " + ncclGetErrorString(r);
  return 1;
}
} // namespace

extern "C" {

static_assert(sizeof(ncclUniqueId) <= EMCNCCL_ID_BYTES, "ncclUniqueId does not fit EMCNCCL_ID_BYTES");

int emcnccl_unique_id(unsigned char id[EMCNCCL_ID_BYTES]) {
  ncclUniqueId u;
  ncclResult_t r = ncclGetUniqueId(&u);
  if (r != ncclSuccess) return failNccl("ncclGetUniqueId", r);
  memset(id, 0, EMCNCCL_ID_BYTES);
  memcpy(id, &u, sizeof u);
  return 0;
}

int emcnccl_init(const unsigned char id[EMCNCCL_ID_BYTES], int rank, int world, int cudaDevice, emcnccl_comm **out) {
  if (!id || !out || rank < 0 || rank >= world) {
    g_error = "emcnccl_init: bad arguments";
    return 1;
  }
  cudaError_t ce = cudaSetDevice(cudaDevice);
  if (ce != cudaSuccess) {
    g_error = std::string("cudaSetDevice: ") + cudaGetErrorString(ce);
    return 1;
  }
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  emcnccl_comm *c = new emcnccl_comm();
  c->rank = rank;
  c->world = world;
  c->device = cudaDevice;
  ncclResult_t r = ncclCommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) {
    delete c;
    return failNccl("ncclCommInitRank", r);
  }
  *out = c;
  return 0;
}

void emcnccl_allreduce_sum_f64(void *user, double *deviceBuffer, int64_t count, void *cudaStream) {
  emcnccl_comm *c = static_cast<emcnccl_comm *>(user);
  ncclResult_t r = ncclAllReduce(deviceBuffer, deviceBuffer, (size_t)count, ncclDouble, ncclSum, c->comm,
                                 static_cast<cudaStream_t>(cudaStream));
  if (r != ncclSuccess) {
    failNccl("ncclAllReduce", r);
    std::fprintf(stderr, "emcnccl: %s\n", g_error.c_str());
  }
  c->calls++;
  c->bytes += count * (int64_t)sizeof(double);
}

int emcnccl_allreduce_sum_host_f64(emcnccl_comm *c, double *host, int64_t count) {
  if (!c || !host || count < 0) {
    g_error = "emcnccl_allreduce_sum_host_f64: bad arguments";
    return 1;
  }
  if (count == 0) return 0;
  cudaSetDevice(c->device);
  if (count > c->stagingCount) {
    if (c->staging) cudaFree(c->staging);
    c->staging = nullptr;
    c->stagingCount = 0;
    if (cudaMalloc(&c->staging, (size_t)count * sizeof(double)) != cudaSuccess) {
      g_error = "emcnccl: cannot allocate the staging buffer";
      return 1;
    }
    c->stagingCount = count;
  }
  cudaError_t e = cudaMemcpy(c->staging, host, (size_t)count * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    ncclResult_t r = ncclAllReduce(c->staging, c->staging, (size_t)count, ncclDouble, ncclSum, c->comm, nullptr);
    if (r != ncclSuccess) return failNccl("ncclAllReduce", r);
    e = cudaStreamSynchronize(nullptr);
  }
  if (e == cudaSuccess) e = cudaMemcpy(host, c->staging, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    g_error = std::string("emcnccl host all-reduce: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

int emcnccl_init_from_file(const char *path, int rank, int world, int cudaDevice, double timeoutSeconds, emcnccl_comm **out) {
  if (!path || !*path) {
    g_error = "emcnccl_init_from_file: no path";
    return 1;
  }
  unsigned char id[EMCNCCL_ID_BYTES];
  if (rank == 0) {
    if (int r = emcnccl_unique_id(id)) return r;
    const std::string tmp = std::string(path) + ".tmp";
    {
      std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
      f.write(reinterpret_cast<const char *>(id), EMCNCCL_ID_BYTES);
      if (!f) {
        g_error = "emcnccl_init_from_file: cannot write " + tmp;
        return 1;
      }
    }
    if (std::rename(tmp.c_str(), path) != 0) {
      g_error = std::string("emcnccl_init_from_file: cannot rename to ") + path;
      return 1;
    }
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      std::ifstream f(path, std::ios::binary);
      if (f && f.read(reinterpret_cast<char *>(id), EMCNCCL_ID_BYTES)) break;
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeoutSeconds) {
        g_error = std::string("emcnccl_init_from_file: no unique id at ") + path;
        return 1;
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
  }
  return emcnccl_init(id, rank, world, cudaDevice, out);
}

int emcnccl_rank(const emcnccl_comm *comm) { return comm ? comm->rank : 0; }
int emcnccl_world(const emcnccl_comm *comm) { return comm ? comm->world : 1; }
int64_t emcnccl_calls(const emcnccl_comm *comm) { return comm ? comm->calls : 0; }
int64_t emcnccl_bytes(const emcnccl_comm *comm) { return comm ? comm->bytes : 0; }

void emcnccl_destroy(emcnccl_comm *comm) {
  if (!comm) return;
  if (comm->staging) cudaFree(comm->staging);
  if (comm->comm) ncclCommDestroy(comm->comm);
  delete comm;
}

const char *emcnccl_last_error(void) { return g_error.c_str(); }

} // extern "C"
