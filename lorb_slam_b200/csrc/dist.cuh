// dist.cuh — NCCL plumbing used by the sharded solves (dist.cu).
#pragma once
#include "common.cuh"
namespace lorb {
bool dist_ready(lorb_ctx* c);
void dist_destroy(lorb_ctx* c);
// in-place sum all-reduce of `n` doubles of DEVICE memory on the ctx stream
int dist_allreduce_sum(lorb_ctx* c, double* dev, size_t n);
// in-place max all-reduce on the bit patterns of non-negative doubles (device)
int dist_allreduce_max_u64(lorb_ctx* c, double* dev, size_t n);
// in-place max all-reduce of `n` ints (device): "is any rank still iterating"
int dist_allreduce_max_i32(lorb_ctx* c, int* dev, size_t n);
int dist_rank(lorb_ctx* c);
int dist_world(lorb_ctx* c);
}  // namespace lorb
