/* libemcnccl.so -- NCCL binding of the sharded device run (include/emcgpu.h: emcgpu_device_set_sharding).
 *
 * The step library (libemcgpu.so) does not link NCCL: it sums its per-step exchange buffers over the ranks through a
 * callback (emcgpu_allreduce_fn).  This small library provides that callback as ncclAllReduce(sum, fp64) over NVLink /
 * NVSwitch, plus the three calls a host needs to form the communicator -- for C++ hosts (the drop-in
 * emcBasicParticleHandler::setSharding) and for the Python bench alike.  One process per GPU.
 *
 * Replaces nothing in the reference (ViennaEMC is single-node OpenMP; SURVEY.md 8e describes the exchange). */
#ifndef EMCNCCL_H
#define EMCNCCL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define EMCNCCL_ID_BYTES 128
typedef struct emcnccl_comm emcnccl_comm;

/* rank 0: a fresh unique id, to be sent to the other ranks by any host-side channel (MPI, torch.distributed, a file) */
int emcnccl_unique_id(unsigned char id[EMCNCCL_ID_BYTES]);
/* every rank: join the communicator (the CUDA device must be the rank's GPU: cudaSetDevice(device) is called) */
int emcnccl_init(const unsigned char id[EMCNCCL_ID_BYTES], int rank, int world, int cudaDevice, emcnccl_comm **out);
/* the emcgpu_allreduce_fn: in-place sum over the ranks of `count` doubles in device memory, on `cudaStream`;
 * user = the emcnccl_comm */
void emcnccl_allreduce_sum_f64(void *user, double *deviceBuffer, int64_t count, void *cudaStream);
/* the same for a small HOST array (staged through a device buffer of the communicator; synchronous): contact counters,
 * ensemble sizes, checksums */
int emcnccl_allreduce_sum_host_f64(emcnccl_comm *comm, double *host, int64_t count);
/* file-based rendezvous for hosts without MPI / torch.distributed: rank 0 creates the unique id and writes it to `path`
 * (atomically), the other ranks wait for the file (at most timeoutSeconds); then every rank joins as in emcnccl_init */
int emcnccl_init_from_file(const char *path, int rank, int world, int cudaDevice, double timeoutSeconds, emcnccl_comm **out);
int emcnccl_rank(const emcnccl_comm *comm);
int emcnccl_world(const emcnccl_comm *comm);
/* number of all-reduce calls issued through this communicator so far / bytes reduced */
int64_t emcnccl_calls(const emcnccl_comm *comm);
int64_t emcnccl_bytes(const emcnccl_comm *comm);
void emcnccl_destroy(emcnccl_comm *comm);
const char *emcnccl_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
