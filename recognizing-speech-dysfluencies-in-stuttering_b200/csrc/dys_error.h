// Status codes + thread-local error text behind dys_last_error().
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <string>

namespace dys {

void set_error(const std::string& msg);
const char* last_error_cstr();

#define DYS_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::dys::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            return DYS_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

}  // namespace dys
