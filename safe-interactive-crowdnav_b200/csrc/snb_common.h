// csrc/snb_common.h -- shared host-side helpers of libsnb.so (error reporting, launch counter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "snb.h"

void snb_set_error(const char *fmt, ...);
void snb_count_launch(int n = 1);

#define SNB_CUDA_TRY(expr)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            snb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
            return SNB_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)

#define SNB_REQUIRE(cond, code, ...)  \
    do {                              \
        if (!(cond)) {                \
            snb_set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)
