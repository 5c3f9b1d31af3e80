// Helpers shared between the CUDA shim (csrc/chain_api.cu) and the host backend (host/chain_backend.cpp).
#pragma once
#include <cuda_runtime_api.h>
#include <stdint.h>

namespace mm2b {
void set_error(const char *fmt, const char *a, const char *b);   // thread-local text behind mm2b_last_error()
bool cuda_ok(cudaError_t e, const char *what);
void count_launches(int n);
long pin_pool_misses();                                          // pinned-pool requests that needed a cudaHostAlloc so far
// host_mapped[0..2] = *a, *b, *c by a one-thread kernel (null: left alone); see seed_kernels.cu
int launch_export_scalars(int64_t *host_mapped, const int64_t *a, const int64_t *b, const int *c, cudaStream_t stream);
}
struct mm2b_workspace;
// device address of a workspace's counters: [0] chunks issued, [1] reads on the general path, [2] reference-semantics cells
extern "C" const unsigned long long *mm2b_ws_counters_dev(const struct mm2b_workspace *ws);
