// Internal helpers shared by the translation units of libabt_b200.so.
#pragma once
#include "../../include/abt_b200.h"

#include <cuda_runtime.h>

namespace abt {
// printf-style; stores a thread-local message and returns `code`
int set_error(int code, const char* fmt, ...);
// 0 when the current device is compute capability 10.x, else ABT_ERR_DEVICE / ABT_ERR_CUDA
int check_device_sm100();
// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
void count_launch(int n = 1);
// blocks (of 32 warps) of the pinned-host span gather: it is confined to this many SMs (abt_debug_set key 15)
extern int g_gather_blocks;
}  // namespace abt
