// Shared host-side helpers for libt2h.so: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/t2h.h"

namespace t2h {

// printf-style; stores the message for t2h_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);
int num_sms();

#define T2H_CHECK_ARG(cond, ...)                        \
  do {                                                  \
    if (!(cond)) return ::t2h::fail(T2H_EINVAL, __VA_ARGS__); \
  } while (0)

#define T2H_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return ::t2h::fail(T2H_ECUDA, "%s failed: %s (%s:%d)", #expr,                  \
                         cudaGetErrorString(e__), __FILE__, __LINE__);              \
  } while (0)

// launch-error check that does not synchronise
#define T2H_LAUNCH_OK() T2H_CUDA(cudaPeekAtLastError())

static inline cudaStream_t as_stream(t2h_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace t2h
