// Helpers shared between the CUDA shim (csrc/chain_api.cu) and the host backend (host/chain_backend.cpp).
#pragma once
#include <cuda_runtime_api.h>

namespace mm2b {
void set_error(const char *fmt, const char *a, const char *b);   // thread-local text behind mm2b_last_error()
bool cuda_ok(cudaError_t e, const char *what);
void count_launches(int n);
}
