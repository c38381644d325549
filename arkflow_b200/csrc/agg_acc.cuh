// agg_acc.cuh — accumulator primitives shared by the tiled and the partitioned GROUP BY kernels.
#pragma once
#include "hash_agg.cuh"
#include "vm.cuh"

namespace ark {

static __device__ __forceinline__ unsigned long long acc_identity(int kind) {
  if (kind == ACC_MIN_I64 || kind == ACC_MIN_F64) return 0x7FFFFFFFFFFFFFFFull;
  if (kind == ACC_MAX_I64 || kind == ACC_MAX_F64) return 0x8000000000000000ull;
  return 0;
}

// dst may be a global or a shared address (generic atomics)
static __device__ __forceinline__ void accumulate(int kind, int arg_is_f64, unsigned long long* dst, unsigned long long bits) {
  switch (kind) {
    case ACC_COUNT_STAR: case ACC_COUNT: atomicAdd(dst, 1ull); break;
    case ACC_SUM_I64: atomicAdd(dst, bits); break;
    case ACC_SUM_F64: atomicAdd(reinterpret_cast<double*>(dst), arg_is_f64 ? __longlong_as_double((long long)bits) : (double)(long long)bits); break;
    case ACC_MIN_I64: atomicMin(reinterpret_cast<long long*>(dst), (long long)bits); break;
    case ACC_MIN_F64: atomicMin(reinterpret_cast<long long*>(dst), f64_total_key(bits)); break;
    case ACC_MAX_I64: atomicMax(reinterpret_cast<long long*>(dst), (long long)bits); break;
    default: atomicMax(reinterpret_cast<long long*>(dst), f64_total_key(bits)); break;
  }
}

// merge a privatised (shared-memory) accumulator into the table
static __device__ __forceinline__ void merge_acc(int kind, unsigned long long* dst, unsigned long long v) {
  switch (kind) {
    case ACC_COUNT_STAR: case ACC_COUNT: case ACC_SUM_I64: atomicAdd(dst, v); break;
    case ACC_SUM_F64: atomicAdd(reinterpret_cast<double*>(dst), __longlong_as_double((long long)v)); break;
    case ACC_MIN_I64: case ACC_MIN_F64: atomicMin(reinterpret_cast<long long*>(dst), (long long)v); break;
    default: atomicMax(reinterpret_cast<long long*>(dst), (long long)v); break;
  }
}

static __device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static __device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static __device__ __forceinline__ long long warp_min_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
  return v;
}
static __device__ __forceinline__ long long warp_max_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}


}  // namespace ark
