// sbx_comm.h -- latitude-band halo exchange between ranks (one process per GPU) over NCCL
// send/recv.  NCCL is resolved at run time with dlopen("libnccl.so.2") so that libsbx.so loads on a
// box without NCCL and shares torch's copy when torch is already in the process.  Internal.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct sbx_comm;

// all return 0 on success; on failure *err points at a static/thread-local message
int sbx_comm_unique_id_impl(uint8_t out128[128], const char **err);
int sbx_comm_init_impl(const uint8_t id128[128], int nranks, int rank, int device, sbx_comm **out, const char **err);
void sbx_comm_destroy_impl(sbx_comm *c);
int sbx_comm_rank(const sbx_comm *c);
int sbx_comm_size(const sbx_comm *c);
// one grouped exchange on `stream`: first_row -> rank-1, last_row -> rank+1, ghosts filled from them
int sbx_comm_halo(sbx_comm *c, const void *first_row, void *top_ghost, const void *last_row, void *bottom_ghost,
                  size_t bytes, cudaStream_t stream, const char **err);
// in-place max-allreduce of two doubles (host values, staged through scratch_dev)
int sbx_comm_allreduce_max2(sbx_comm *c, double v[2], double *scratch_dev, cudaStream_t stream, const char **err);
