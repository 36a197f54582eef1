/*
 * mgic_comm.h -- multi-GPU plumbing of the z-slab decomposition (one process per GPU).
 *
 * Replaces, for this path, what the reference reaches through Chombo: LevelData::exchange with the face-only
 * copier (Source/VariableCoeffPoissonOperatorFactory.cpp:82-99; call sites VariableCoeffPoissonOperator.cpp:48,131,
 * 163,301,384) and the MPI_Allreduce inside norm / dotProduct.  Every MG level's domain is cut into z-slabs, rank r
 * owning global planes [k0, k0 + nz_local); halo planes are contiguous nx*ny doubles, exchanged with the two
 * z-neighbours over NVLink on the context's stream -- by peer stores into CUDA-IPC-mapped ghost planes (one kernel per
 * exchange, flags in peer memory), or by ncclSend / ncclRecv where the mapping is unavailable; scalars by ncclAllReduce.
 *
 * The library dlopen()s libnccl.so.2 (the copy torch already loaded); nothing here is needed on one GPU.
 */
#ifndef MGIC_COMM_H
#define MGIC_COMM_H

#include "mgic.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MGIC_NCCL_ID_BYTES 128

/* rank 0: create the NCCL unique id; the caller broadcasts the 128 bytes to all ranks (torch.distributed / MPI) */
int mgic_comm_unique_id(unsigned char id[MGIC_NCCL_ID_BYTES]);
/* every rank: join the communicator; installs the halo-exchange and all-reduce hooks on the context */
int mgic_comm_init(mgic_ctx *, const unsigned char id[MGIC_NCCL_ID_BYTES], int rank, int nranks);
int mgic_comm_destroy(mgic_ctx *);
/* exchange `planes` (1 or 2) ghost planes of a field with the z-neighbours (periodic wrap if the level is periodic) */
int mgic_comm_halo_exchange(mgic_ctx *, mgic_field *, int planes);
/* bytes this rank has sent through halo exchanges since mgic_comm_init (bench reporting) */
long long mgic_comm_halo_bytes(mgic_ctx *);
/* how many exchanges went through NVLink peer stores (k_halo_push over CUDA IPC mappings) and how many through
 * ncclSend/ncclRecv since mgic_comm_init; p2p_available = the ranks could map each other's control block.
 * Peer stores are the default; MGIC_P2P_HALO=0 in the environment or mgic_ctx_set_option("p2p_halo", 0) (same value
 * on every rank) selects NCCL. */
int mgic_comm_halo_stats(mgic_ctx *, long long *p2p_exchanges, long long *nccl_exchanges, int *p2p_available);

#ifdef __cplusplus
}
#endif
#endif
