// launch.h - host entry points of the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include "plan.h"

namespace racu {
cudaError_t launch_plan(Plan const & plan, cudaStream_t stream);
cudaError_t launch_static(Plan const & plan, cudaStream_t stream);
cudaError_t launch_jit(Plan const & plan, cudaStream_t stream);
int jit_occupancy(Plan const & plan);      // resident CTAs per SM of the plan's run-time specialised kernel (0: unknown)
bool try_transpose(Plan const & plan, cudaStream_t stream, cudaError_t & err);
cudaError_t launch_fill(void * dst, int dtype, racu_base v, size_t n, cudaStream_t stream);
int64_t launch_count();
void launch_count_reset();
void launch_count_add(int64_t n);
} // namespace racu
