// Multi-GPU plumbing of the query path: queries are independent units, the index image
// is replicated, and the ONLY exchange is one ncclAllGather of every rank's packed per-query result
// block, issued on the batch's own stream right behind its last kernel.  NCCL is bound at run time
// (dlopen of libnccl.so.2: inside a torch process that resolves to the NCCL torch already loaded, so
// one NCCL serves the process; a plain C/Rust caller gets the system library) — the product library
// itself has no link-time dependency on it and single-GPU callers never touch it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

struct pb_comm;

namespace pbg {

// one ncclAllGather of `bytes` per rank (send -> recv[rank * bytes]) on `st`
int comm_allgather(pb_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t st);
int comm_world(const pb_comm* c);
int comm_rank(const pb_comm* c);
int comm_device(const pb_comm* c);
int nccl_group_start();
int nccl_group_end();

}  // namespace pbg
