// stage_store.cuh — write a CTA's contiguous output byte range from shared memory to global memory with
// destination-aligned 16-byte stores.  The staging buffer must start at the destination's misalignment
// (stage_misalignment) so that shared and global addresses agree mod 16.
#pragma once
#include <cstdint>

namespace ark {

__device__ __forceinline__ int stage_misalignment(const uint8_t* gdst) { return (int)(reinterpret_cast<uintptr_t>(gdst) & 15); }

// `stage` is 16-byte aligned; the range's bytes sit at stage[mis .. mis + bytes).  Call from every thread
// of the CTA after a __syncthreads().
__device__ __forceinline__ void stage_store(uint8_t* gdst, const uint8_t* stage, int mis, int bytes, int tid, int nthreads) {
  const int head = min((16 - mis) & 15, bytes);
  if (tid < head) gdst[tid] = stage[mis + tid];
  const int body = (bytes - head) >> 4;
  const uint4* sv = reinterpret_cast<const uint4*>(stage + mis + head);  // mis + head ≡ 0 (mod 16) whenever body > 0
  uint4* gv = reinterpret_cast<uint4*>(gdst + head);
  for (int g = tid; g < body; g += nthreads) gv[g] = sv[g];
  const int tail0 = head + body * 16;
  if (tid < bytes - tail0) gdst[tail0 + tid] = stage[mis + tail0 + tid];
}

}  // namespace ark
