// filter_project_tma.cu — the fast path of filter + project + compaction for the common shape
//   SELECT <≤2 fixed-width columns> [, <one Utf8/Binary column>] FROM t WHERE <col> <cmp> <literal>
// (BASELINE config 2: SELECT sensor, value FROM flow WHERE value >= 10).
//
// Same single-pass algorithm as filter_project.cu (ticketed tiles, ballot ranking, decoupled
// look-back, shared-memory staged coalesced stores) but restructured for Blackwell's async copy path:
//   * persistent CTAs (2 per SM), one PRODUCER warp + 8 CONSUMER warps (warp specialisation);
//   * the producer takes a tile ticket and issues 1-D TMA bulk copies (cp.async.bulk → UBLKCP) of the
//     tile's value columns, offsets and string bytes into a 2-stage shared-memory ring, completion
//     tracked by mbarriers (expect_tx); HBM latency is hidden by the ring, not by occupancy;
//   * consumers read everything from shared memory: predicate, ranks, word-granular string compaction
//     (aligned source words funnel-shifted into destination words), then vectorised global stores.
// Bulk copies only ever touch 16-byte blocks that contain at least one valid byte of the source
// buffer, so they never cross into an unmapped page (blocks do not straddle pages).
#include "batch.h"
#include "filter_project.cuh"
#include "vm.cuh"

namespace ark {

namespace {

constexpr int TT = 1024;                 // rows per tile
constexpr int T_CONSUMERS = 256;         // 8 consumer warps
constexpr int T_THREADS = T_CONSUMERS + 32;
constexpr int T_CHUNKS = TT / T_CONSUMERS;  // 4 rows per consumer thread, interleaved by 256
constexpr int T_WARPS = T_CONSUMERS / 32;
constexpr int T_STAGES = 2;
constexpr int T_MAX_FIXED = 2;           // fixed-width columns staged per tile (predicate column first)
constexpr int T_STR_CAP = 16 * 1024;     // staged string bytes per tile; larger tiles use the global path

struct TmaParams {
  int64_t n_rows;
  int32_t n_tiles;
  int32_t n_fixed;                 // staged fixed-width columns; [0] is the predicate column
  int32_t has_varlen;
  int32_t sp_cmp, sp_is_f64;
  uint64_t sp_const;
  const unsigned long long* fixed_in[T_MAX_FIXED];
  unsigned long long* fixed_out[T_MAX_FIXED];   // nullptr ⇒ staged for the predicate only
  const int32_t* offsets_in;
  const uint8_t* data_in;
  int32_t* offsets_out;
  uint8_t* data_out;
  unsigned long long* desc;
  unsigned int* ticket;
  long long* totals;
};

struct __align__(16) Stage {
  unsigned long long fixed[T_MAX_FIXED][TT];   // 16 KB
  int32_t offsets[TT + 4];                     // rows+1 offsets, padded to a 16-byte multiple
  uint8_t bytes[T_STR_CAP + 32];               // [align-down(S), align-up(E)) window of the string bytes
};

struct __align__(16) OutStage {
  unsigned long long fixed[T_MAX_FIXED][TT];
  int32_t offsets[TT];
  uint8_t bytes[T_STR_CAP + 32];
};

struct StageMeta {
  int32_t tile;       // -1 ⇒ no more tiles
  int32_t rows;
  int32_t str_base;   // offsets value that maps to bytes[0] (= align-down of S in absolute bytes, relative to data_in)
  int32_t str_staged; // 1 ⇒ string bytes are in shared memory
};

struct __align__(16) Smem {
  Stage in[T_STAGES];
  OutStage out;
  unsigned long long full_bar[T_STAGES];
  unsigned long long empty_bar[T_STAGES];
  StageMeta meta[T_STAGES];
  int s_cnt[T_CHUNKS * T_WARPS];
  int s_bytes[T_CHUNKS * T_WARPS];
  long long s_excl[2];
  int s_total[2];
};

constexpr unsigned long long DESC_AGG = 1ull << 62;
constexpr unsigned long long DESC_PREFIX = 2ull << 62;
constexpr unsigned long long DESC_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\tbra.uni WAIT_LOOP;\n\tWAIT_DONE:\n\t}"
      ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(T_CONSUMERS) : "memory"); }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ long long lookback(unsigned long long* desc, int tile, int ch, long long agg, int lane) {
  unsigned long long* mine = desc + (size_t)tile * FP_CHANNELS + ch;
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(mine, DESC_PREFIX | (unsigned long long)agg);
    return 0;
  }
  if (lane == 0) st_volatile_u64(mine, DESC_AGG | (unsigned long long)agg);
  long long running = 0;
  int look = tile - 1;
  while (true) {
    const int idx = look - lane;
    unsigned long long d = DESC_PREFIX;
    if (idx >= 0) {
      do { d = ld_volatile_u64(desc + (size_t)idx * FP_CHANNELS + ch); } while ((d >> 62) == 0);
    }
    __syncwarp();
    const unsigned pm = __ballot_sync(0xffffffffu, (d >> 62) == 2);
    long long val = (long long)(d & DESC_MASK);
    if (pm) {
      const int first = __ffs(pm) - 1;
      if (lane > first) val = 0;
      running += warp_sum(val);
      break;
    }
    running += warp_sum(val);
    look -= 32;
  }
  if (lane == 0) st_volatile_u64(mine, DESC_PREFIX | (unsigned long long)(running + agg));
  return running;
}

// copy len bytes inside shared memory, word-granular on the destination
__device__ __forceinline__ void smem_copy(uint8_t* dst, const uint8_t* src, int len) {
  const unsigned d0 = smem_addr(dst), s0 = smem_addr(src);
  if (((d0 | s0 | (unsigned)len) & 3) == 0) {  // all word aligned (fixed-length keys such as "temp_0000123")
    const unsigned* s = reinterpret_cast<const unsigned*>(src);
    unsigned* d = reinterpret_cast<unsigned*>(dst);
    for (int i = 0; i < (len >> 2); ++i) d[i] = s[i];
    return;
  }
  int i = 0;
  for (; i < len && ((d0 + i) & 3); ++i) dst[i] = src[i];           // head: up to 3 bytes
  const int words = (len - i) >> 2;
  if (words > 0) {
    const unsigned sa = s0 + i;
    const unsigned* sw = reinterpret_cast<const unsigned*>(src + i - (sa & 3));  // aligned word containing src[i]
    const unsigned sh = (sa & 3) * 8;
    unsigned* d = reinterpret_cast<unsigned*>(dst + i);
    unsigned lo = sw[0];
    for (int w = 0; w < words; ++w) {
      const unsigned hi = sh ? sw[w + 1] : 0;
      d[w] = sh ? __funnelshift_r(lo, hi, sh) : lo;
      lo = sh ? hi : sw[w + 1 < words ? w + 1 : w];
    }
    i += words * 4;
  }
  for (; i < len; ++i) dst[i] = src[i];                              // tail
}

__global__ void __launch_bounds__(T_THREADS, 2) filter_project_tma_kernel(const __grid_constant__ TmaParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < T_STAGES; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.empty_bar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= T_CONSUMERS) {
    // ===================== producer warp =====================
    if (tid != T_CONSUMERS) return;
    for (int it = 0;; ++it) {
      const int s = it % T_STAGES;
      if (it >= T_STAGES) mbar_wait(&S.empty_bar[s], ((it / T_STAGES) - 1) & 1);
      const int tile = (int)atomicAdd(P.ticket, 1u);
      StageMeta m;
      if (tile >= P.n_tiles) {
        m.tile = -1; m.rows = 0; m.str_base = 0; m.str_staged = 0;
        S.meta[s] = m;
        mbar_arrive(&S.full_bar[s]);
        break;
      }
      const int64_t row0 = (int64_t)tile * TT;
      const int rows = (int)((P.n_rows - row0) < TT ? (P.n_rows - row0) : TT);
      unsigned tx = 0;
      const unsigned fixed_bytes = ((unsigned)rows * 8 + 15) & ~15u;
      const unsigned off_bytes = ((unsigned)(rows + 1) * 4 + 15) & ~15u;
      unsigned str_bytes = 0;
      const uint8_t* str_src = nullptr;
      m.tile = tile; m.rows = rows; m.str_base = 0; m.str_staged = 0;
      if (P.has_varlen) {
        const int32_t o0 = P.offsets_in[row0], o1 = P.offsets_in[row0 + rows];
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(P.data_in + o0), a1 = reinterpret_cast<uintptr_t>(P.data_in + o1);
        const uintptr_t lo = a0 & ~(uintptr_t)15, hi = (a1 + 15) & ~(uintptr_t)15;
        m.str_base = o0 - (int32_t)(a0 - lo);
        if (o1 > o0 && hi - lo <= (uintptr_t)T_STR_CAP + 16) { str_bytes = (unsigned)(hi - lo); str_src = reinterpret_cast<const uint8_t*>(lo); m.str_staged = 1; }
        else if (o1 == o0) m.str_staged = 1;
        tx += off_bytes;
      }
      tx += (unsigned)P.n_fixed * fixed_bytes + str_bytes;
      S.meta[s] = m;
      mbar_expect_tx(&S.full_bar[s], tx);
      for (int c = 0; c < P.n_fixed; ++c) tma_load_1d(S.in[s].fixed[c], P.fixed_in[c] + row0, fixed_bytes, &S.full_bar[s]);
      if (P.has_varlen) {
        tma_load_1d(S.in[s].offsets, P.offsets_in + row0, off_bytes, &S.full_bar[s]);
        if (str_bytes) tma_load_1d(S.in[s].bytes, str_src, str_bytes, &S.full_bar[s]);
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1;
  const long long kc = P.sp_is_f64 ? f64_total_key(P.sp_const) : (long long)P.sp_const;
  for (int it = 0;; ++it) {
    const int s = it % T_STAGES;
    mbar_wait(&S.full_bar[s], (it / T_STAGES) & 1);
    const StageMeta m = S.meta[s];
    if (m.tile < 0) break;
    const Stage& in = S.in[s];
    const int tile = m.tile, rows = m.rows;
    const int64_t row0 = (int64_t)tile * TT;

    // ---- A: predicate, per-warp counts, selected string lengths ----
    unsigned flags = 0;
    unsigned bal[T_CHUNKS];
    int bp[T_CHUNKS];
#pragma unroll
    for (int k = 0; k < T_CHUNKS; ++k) {
      const int lr = k * T_CONSUMERS + tid;
      bool f = lr < rows;
      if (f) {
        const unsigned long long v = in.fixed[0][lr];
        f = cmp_i64(P.sp_cmp, P.sp_is_f64 ? f64_total_key(v) : (long long)v, kc);
      }
      bal[k] = __ballot_sync(0xffffffffu, f);
      flags |= (unsigned)f << k;
      if (lane == 0) S.s_cnt[k * T_WARPS + warp] = __popc(bal[k]);
      if (P.has_varlen) {
        int len = 0;
        if (f) len = in.offsets[lr + 1] - in.offsets[lr];
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        bp[k] = incl - len;
        if (lane == 31) S.s_bytes[k * T_WARPS + warp] = incl;
      }
    }
    consumer_sync();
    // ---- B: tile scan + decoupled look-back (warp 0) ----
    if (warp == 0) {
      {
        const int c = S.s_cnt[lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        S.s_cnt[lane] = incl - c;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const long long ex = lookback(P.desc, tile, 0, total, lane);
        if (lane == 0) { S.s_excl[0] = ex; S.s_total[0] = total; if (tile == P.n_tiles - 1) P.totals[0] = ex + total; }
      }
      if (P.has_varlen) {
        const int c = S.s_bytes[lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        S.s_bytes[lane] = incl - c;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const long long ex = lookback(P.desc, tile, 1, total, lane);
        if (lane == 0) { S.s_excl[1] = ex; S.s_total[1] = total; if (tile == P.n_tiles - 1) P.totals[1] = ex + total; }
      }
    }
    consumer_sync();
    const int tile_cnt = S.s_total[0];
    const long long base_cnt = S.s_excl[0];
    const int tb = P.has_varlen ? S.s_total[1] : 0;
    const long long bb = P.has_varlen ? S.s_excl[1] : 0;
    const int shift = (int)(bb & 15);
    const bool str_fast = P.has_varlen && m.str_staged && tb <= T_STR_CAP;
    // ---- C: stage the surviving rows ----
#pragma unroll
    for (int k = 0; k < T_CHUNKS; ++k) {
      if (!((flags >> k) & 1)) continue;
      const int lr = k * T_CONSUMERS + tid;
      const int rank = S.s_cnt[k * T_WARPS + warp] + __popc(bal[k] & lt_mask);
      for (int c = 0; c < P.n_fixed; ++c) if (P.fixed_out[c]) S.out.fixed[c][rank] = in.fixed[c][lr];
      if (P.has_varlen) {
        const int lbp = S.s_bytes[k * T_WARPS + warp] + bp[k];
        S.out.offsets[rank] = (int32_t)(bb + lbp);
        const int o0 = in.offsets[lr], len = in.offsets[lr + 1] - o0;
        if (str_fast) smem_copy(S.out.bytes + shift + lbp, in.bytes + (o0 - m.str_base), len);
        else {  // long strings: straight from global to global
          const uint8_t* src = P.data_in + o0;
          uint8_t* dst = P.data_out + bb + lbp;
          for (int i = 0; i < len; ++i) dst[i] = src[i];
        }
      }
    }
    consumer_sync();
    if (tid == 0) mbar_arrive(&S.empty_bar[s]);  // the input stage is free: the producer may refill it
    // ---- D: coalesced stores ----
    for (int c = 0; c < P.n_fixed; ++c) {
      if (!P.fixed_out[c]) continue;
      unsigned long long* dst = P.fixed_out[c] + base_cnt;
      for (int i = tid; i < tile_cnt; i += T_CONSUMERS) dst[i] = S.out.fixed[c][i];
    }
    if (P.has_varlen) {
      for (int i = tid; i < tile_cnt; i += T_CONSUMERS) P.offsets_out[base_cnt + i] = S.out.offsets[i];
      if (tile == P.n_tiles - 1 && tid == 0) P.offsets_out[base_cnt + tile_cnt] = (int32_t)(bb + tb);
      if (str_fast) {
        uint8_t* gbase = P.data_out + (bb - shift);
        const int total = shift + tb;
        for (int p = tid * 16; p < total; p += T_CONSUMERS * 16) {
          if (p >= shift && p + 16 <= total) *reinterpret_cast<uint4*>(gbase + p) = *reinterpret_cast<const uint4*>(S.out.bytes + p);
          else {
            const int q0 = p > shift ? p : shift, q1 = (p + 16 < total) ? p + 16 : total;
            for (int q = q0; q < q1; ++q) gbase[q] = S.out.bytes[q];
          }
        }
      }
    }
    // no trailing barrier: the next tile's first consumer_sync() orders these reads of S.out / s_* against its writes
  }
}

}  // namespace

struct TmaLaunchArgs {
  TmaParams P;
};

size_t filter_project_tma_smem_bytes() { return sizeof(Smem) + 128; }

// Returns false when the inputs do not meet the alignment rules of the bulk-copy path.
bool launch_filter_project_tma(int64_t n_rows, int n_fixed, const void* const* fixed_in, void* const* fixed_out, const int32_t* offsets_in,
                               const uint8_t* data_in, int32_t* offsets_out, uint8_t* data_out, int cmp, int is_f64, uint64_t constant,
                               unsigned long long* desc, unsigned int* ticket, long long* totals, cudaStream_t stream) {
  for (int c = 0; c < n_fixed; ++c) if (reinterpret_cast<uintptr_t>(fixed_in[c]) & 15) return false;
  if (offsets_in && (reinterpret_cast<uintptr_t>(offsets_in) & 15)) return false;
  if (data_out && (reinterpret_cast<uintptr_t>(data_out) & 15)) return false;
  static bool configured = false;
  static int num_sms = 148;
  if (!configured) {
    ARK_CUDA(cudaFuncSetAttribute(filter_project_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)filter_project_tma_smem_bytes()));
    int dev = 0;
    ARK_CUDA(cudaGetDevice(&dev));
    ARK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  TmaParams P;
  memset(&P, 0, sizeof P);
  P.n_rows = n_rows; P.n_tiles = (int)ceil_div(n_rows, TT); P.n_fixed = n_fixed; P.has_varlen = offsets_in != nullptr;
  P.sp_cmp = cmp; P.sp_is_f64 = is_f64; P.sp_const = constant;
  for (int c = 0; c < n_fixed; ++c) { P.fixed_in[c] = (const unsigned long long*)fixed_in[c]; P.fixed_out[c] = (unsigned long long*)fixed_out[c]; }
  P.offsets_in = offsets_in; P.data_in = data_in; P.offsets_out = offsets_out; P.data_out = data_out;
  P.desc = desc; P.ticket = ticket; P.totals = totals;
  const int grid = std::min(P.n_tiles, num_sms * 2);
  KernelTimer t("filter_project_tma_kernel", stream);
  filter_project_tma_kernel<<<grid, T_THREADS, filter_project_tma_smem_bytes(), stream>>>(P);
  return true;
}

int filter_project_tma_tile_rows() { return TT; }

}  // namespace ark
