// host_copy.cpp — pageable → pinned copies for the staging step of batch.cu (compiled by g++, not nvcc: AVX intrinsics).
//
// glibc's memcpy moved 3-4 GB/s per thread on the B200 hosts (2 x Xeon 8562Y+) for the 4 MB chunks staged here — 26 GB/s
// with eight threads, half of what the PCIe link takes (profiles/r2_e2e_staging.txt).  The destination is a pinned slot
// that the CPU never reads again (the DMA engine does), so the copy is written with non-temporal stores: no
// read-for-ownership of the destination lines, no cache pollution, and the loads run a software prefetch ahead.
#include <immintrin.h>

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace ark {

namespace {

__attribute__((target("avx512f"))) void copy_nt512(char* dp, const char* sp, size_t n) {
  size_t i = 0;
  for (; i + 256 <= n; i += 256) {
    const __m512i a = _mm512_loadu_si512(sp + i), b = _mm512_loadu_si512(sp + i + 64);
    const __m512i c = _mm512_loadu_si512(sp + i + 128), d = _mm512_loadu_si512(sp + i + 192);
    _mm_prefetch(sp + i + 2048, _MM_HINT_T0); _mm_prefetch(sp + i + 2112, _MM_HINT_T0);
    _mm_prefetch(sp + i + 2176, _MM_HINT_T0); _mm_prefetch(sp + i + 2240, _MM_HINT_T0);
    _mm512_stream_si512(reinterpret_cast<__m512i*>(dp + i), a); _mm512_stream_si512(reinterpret_cast<__m512i*>(dp + i + 64), b);
    _mm512_stream_si512(reinterpret_cast<__m512i*>(dp + i + 128), c); _mm512_stream_si512(reinterpret_cast<__m512i*>(dp + i + 192), d);
  }
  if (i < n) memcpy(dp + i, sp + i, n - i);
}

__attribute__((target("avx2"))) void copy_nt256(char* dp, const char* sp, size_t n) {
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(sp + i)), b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(sp + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(sp + i + 64)), d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(sp + i + 96));
    _mm_prefetch(sp + i + 1024, _MM_HINT_T0); _mm_prefetch(sp + i + 1088, _MM_HINT_T0);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dp + i), a); _mm256_stream_si256(reinterpret_cast<__m256i*>(dp + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dp + i + 64), c); _mm256_stream_si256(reinterpret_cast<__m256i*>(dp + i + 96), d);
  }
  if (i < n) memcpy(dp + i, sp + i, n - i);
}

int detect() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return 2;
  if (__builtin_cpu_supports("avx2")) return 1;
  return 0;
}

}  // namespace

// kind: 0 memcpy, 1 AVX2 non-temporal, 2 AVX-512 non-temporal, -1 best available (ARK_STAGE_COPY overrides in batch.cu)
int host_copy_stream(void* dst, const void* src, size_t n, int kind) {
  static const int best = detect();
  if (kind < 0 || kind > best) kind = best;
  char* dp = static_cast<char*>(dst);
  const char* sp = static_cast<const char*>(src);
  if (kind == 0 || n < 4096) { memcpy(dp, sp, n); return kind; }
  const size_t head = (64 - (reinterpret_cast<uintptr_t>(dp) & 63)) & 63;  // streaming stores want an aligned destination
  if (head) { memcpy(dp, sp, head); dp += head; sp += head; n -= head; }
  if (kind == 2) copy_nt512(dp, sp, n); else copy_nt256(dp, sp, n);
  _mm_sfence();  // the chunk is handed to the DMA engine next: the streaming stores must be globally visible
  return kind;
}

}  // namespace ark

extern "C" int ark_host_copy(void* dst, const void* src, int64_t n, int kind) {
  if (!dst || !src || n < 0) return -1;
  return ark::host_copy_stream(dst, src, (size_t)n, kind);
}
