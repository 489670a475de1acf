// nann_b200.cu -- single translation unit of libnann_b200.so (sm_100a).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "traverse_kernels.cuh"
#include "scorer_mlp_exact.cuh"

#include "lib_core.inl"
#include "lib_ops.inl"
#include "lib_ragged.inl"
#include "lib_search.inl"
#include "lib_shard.inl"
#include "lib_dist.inl"
#include "builder_kernels.cuh"
#include "lib_builder.inl"
#include "lib_bloom.inl"
#include "lib_eval.inl"
#include "lib_executor.inl"
