// Helpers shared between the CUDA shim (csrc/chain_api.cu) and the host backend (host/chain_backend.cpp).
#pragma once
#include <cuda_runtime_api.h>

namespace mm2b {
void set_error(const char *fmt, const char *a, const char *b);   // thread-local text behind mm2b_last_error()
bool cuda_ok(cudaError_t e, const char *what);
void count_launches(int n);
}
struct mm2b_workspace;
// device address of a workspace's counters: [0] chunks issued, [1] reads on the general path, [2] reference-semantics cells
extern "C" const unsigned long long *mm2b_ws_counters_dev(const struct mm2b_workspace *ws);
