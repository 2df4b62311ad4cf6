// Host-side helpers shared by the C-ABI translation units (error reporting, tensor-map encoding).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stablemtl_sm100.h"

namespace smtl_host {

void set_error(const char* fmt, ...);
const char* get_error();

#define SMTL_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            smtl_host::set_error(__VA_ARGS__); \
            return SMTL_EINVAL;              \
        }                                    \
    } while (0)

#define SMTL_CHECK_CUDA(expr)                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            smtl_host::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SMTL_ECUDA;                                                                  \
        }                                                                                       \
    } while (0)

// Encode a 2-D bf16 row-major tensor map [rows, cols] (leading dim ld elements), box = [box_rows, 64 cols],
// 128-byte swizzle, zero OOB fill.  Returns SMTL_* code.
int encode_tmap_bf16_2d(uint64_t out[16], const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols = 64);

int num_sms();   // of the CURRENT device (cached per device ordinal)

// True the first time it is called for `mask` on the current device (then records the device's bit): per-device,
// thread-safe guard for cudaFuncSetAttribute, which applies to the current device only.
bool first_use_on_device(std::atomic<uint64_t>& mask);

}  // namespace smtl_host
