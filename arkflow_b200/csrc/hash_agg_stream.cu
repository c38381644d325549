// hash_agg_stream.cu — GROUP BY for high-cardinality keys: the row kernel of hash_agg.cu rebuilt as a persistent,
// software-pipelined kernel with several probe chains in flight per thread.
//
// hash_agg_kernel (hash_agg.cu) walks one row per thread through three DEPENDENT memory latencies — offsets → key
// bytes → table slot — and resolves the probe chains of a thread one after the other; ncu (profiles/r1l_hash_agg_ncu.json)
// shows the result: 15 long-scoreboard stalls per issue, issue slots 36 % busy, DRAM 11 %, 0.097 of the HBM roofline at
// 10^6 groups.  Here:
//   * a CTA stays resident and takes 256·R-row tiles; the loads of the NEXT tile (offsets / Int64 keys / predicate
//     column into registers, the tile's key bytes by one 1-D TMA bulk copy into the other half of a two-stage ring)
//     are issued before the current tile is touched, so the only latency on a row's critical path is the table's;
//   * every thread owns R consecutive rows: the home BUCKETS of all R rows (four keys each, hash_agg.cuh) are requested
//     before any is inspected — R independent L2 requests per thread — and a row that is not settled by its home
//     bucket (4 % of the rows at load 0.48) continues with table_find_or_claim;
//   * aggregate arguments are fetched with 16-byte loads while the probes are in flight.
// Table layout, key encoding, claim protocol (128-bit CAS), accumulators (fire-and-forget REDs) and the overflow /
// group-count protocol are exactly hash_agg_kernel's, so everything downstream (compaction, emit, multi-GPU
// partition) is shared.  Algorithmic traffic: key + argument bytes read once (24 B/row, config 3).
#include <atomic>

#include "agg_acc.cuh"
#include "engine.h"
#include "hash_agg.cuh"
#include "hashkey.cuh"
#include "tma.cuh"

namespace ark {

namespace {

constexpr int HS_THREADS = 256;
constexpr int HS_PRE = 2;  // aggregate arguments prefetched into registers per tile (further ones are loaded in place)


template <int R>
__device__ __forceinline__ void ld_vec_u64(const unsigned long long* src, bool fast, int lr0, int rows, unsigned long long (&v)[R]) {
  if constexpr (R == 1) {
    v[0] = 0;
    if (lr0 < rows) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v[0]) : "l"(src));
  } else {
    if (fast) {
#pragma unroll
      for (int j = 0; j < R; j += 2)
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v[j]), "=l"(v[j + 1]) : "l"(src + j));
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) v[j] = (lr0 + j < rows) ? src[j] : 0;
    }
  }
}

template <int R>
__device__ __forceinline__ void ld_vec_off(const int32_t* os, bool fast, int lr0, int rows, int lane, int (&off)[R], int* offx) {
  if (fast) {
    if constexpr (R == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(off[0]), "=r"(off[1]), "=r"(off[2]), "=r"(off[3]) : "l"(os));
    else if constexpr (R == 2) asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(off[0]), "=r"(off[1]) : "l"(os));
    else asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(off[0]) : "l"(os));
  } else {
#pragma unroll
    for (int j = 0; j < R; ++j) off[j] = (lr0 + j <= rows) ? os[j] : 0;
  }
  *offx = ((lane == 31 || lr0 + R >= rows) && lr0 + R <= rows) ? os[R] : 0;
}

// SIG 1: the accumulators are exactly {COUNT(*), SUM(<non-null Int64 column>)} (config 3): two REDs, no dispatch.
template <int KEYK, int PRED, int R, int SIG>
__global__ void __launch_bounds__(HS_THREADS) hash_agg_stream_kernel(const __grid_constant__ AggParams P, const int str_cap, const int dbg) {
  constexpr int TR = HS_THREADS * R;
  constexpr int PRODUCER = HS_THREADS - 32;
  extern __shared__ __align__(16) uint8_t smem[];  // KEY_BYTES: [key bytes stage 0][stage 1], each str_cap + 32
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ int s_str_base[2], s_str_staged[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t n = P.n_rows;
  const int n_tiles = (int)((n + TR - 1) / TR);
  const int lr0 = R * tid;
  const int stage_bytes = str_cap + 32;
  const ColView& kc = P.cols[P.key_slot];
  const unsigned long long bmask = P.mask >> 2;
  const int bstride = P.bucket_stride;
  auto tile_rows = [&](int t) { const int64_t r = n - (int64_t)t * TR; return (int)(r < TR ? r : TR); };
  auto issue_window = [&](int st, int32_t o0, int32_t o1) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>((const uint8_t*)kc.data + o0), a1 = reinterpret_cast<uintptr_t>((const uint8_t*)kc.data + o1);
    const uintptr_t lo = a0 & ~(uintptr_t)15, hi = (a1 + 15) & ~(uintptr_t)15;
    int staged = 0;
    if (o1 > o0 && hi - lo <= (uintptr_t)str_cap) {
      staged = 1;
      mbar_expect_tx(&s_bar[st], (unsigned)(hi - lo));
      tma_load_1d(smem + st * stage_bytes, reinterpret_cast<const void*>(lo), (unsigned)(hi - lo), &s_bar[st]);
    }
    s_str_base[st] = o0 - (int32_t)(a0 - lo); s_str_staged[st] = staged;
  };
  // which accumulators get their argument prefetched (the first HS_PRE that have a column argument)
  int pre_acc[HS_PRE];
  {
    int k = 0;
#pragma unroll
    for (int i = 0; i < HS_PRE; ++i) pre_acc[i] = -1;
    for (int a = 0; a < P.n_acc && k < HS_PRE; ++a)
      if (P.accs[a].kind != ACC_COUNT_STAR) pre_acc[k++] = a;
  }

  int tile = blockIdx.x;
  int32_t bo0 = 0, bo1 = 0;
  if (tid == PRODUCER) {
    if (KEYK == KEY_BYTES) {
      mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init();
      if (tile < n_tiles) { const int64_t r0 = (int64_t)tile * TR; issue_window(0, kc.offsets[r0], kc.offsets[r0 + tile_rows(tile)]); }
      const int t1 = tile + (int)gridDim.x;
      if (t1 < n_tiles) { const int64_t r1 = (int64_t)t1 * TR; bo0 = kc.offsets[r1]; bo1 = kc.offsets[r1 + tile_rows(t1)]; }
    }
  }
  __syncthreads();
  // registers of the next tile
  int offn[R] = {}; int offxn = 0;
  unsigned long long kvn[R] = {}, pvn[R] = {};
  auto prefetch = [&](int t) {
    const int64_t r0 = (int64_t)t * TR;
    const int rows = tile_rows(t);
    const bool full = lr0 + R <= rows;
    if (KEYK == KEY_BYTES) {
      const int32_t* os = kc.offsets + r0 + lr0;
      ld_vec_off<R>(os, full && (reinterpret_cast<uintptr_t>(os) & (R * 4 - 1)) == 0, lr0, rows, lane, offn, &offxn);
    } else {
      const unsigned long long* ks = (const unsigned long long*)kc.data + r0 + lr0;
      ld_vec_u64<R>(ks, full && (reinterpret_cast<uintptr_t>(ks) & 15) == 0, lr0, rows, kvn);
    }
    if (PRED == 1) {
      const unsigned long long* ps = (const unsigned long long*)P.cols[P.sp_slot].data + r0 + lr0;
      ld_vec_u64<R>(ps, full && (reinterpret_cast<uintptr_t>(ps) & 15) == 0, lr0, rows, pvn);
    }
  };
  if (tile < n_tiles) prefetch(tile);

  unsigned int claimed = 0;
  unsigned ph = 0;
  int32_t err_overflow = 0, long_flag = 0;
  const long long pred_c = P.sp_is_f64 ? f64_total_key(P.sp_const) : (long long)P.sp_const;
  for (int it = 0; tile < n_tiles; ++it, tile += (int)gridDim.x) {
    const int st = it & 1;
    const int64_t row0 = (int64_t)tile * TR;
    const int rows = tile_rows(tile);
    const bool full = lr0 + R <= rows;
    int off[R + 1];
    unsigned long long kv[R], pv[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { off[j] = offn[j]; kv[j] = kvn[j]; pv[j] = pvn[j]; }
    const int offx = offxn;
    // ---- A: the next tile's loads ----
    const int next = tile + (int)gridDim.x;
    int stop = 0;
    if (tid == PRODUCER) {
      stop = *reinterpret_cast<volatile int32_t*>(P.overflow);  // table too small: the host retries with 4× the slots
      if (KEYK == KEY_BYTES) {
        if (next < n_tiles) issue_window(st ^ 1, bo0, bo1);
        const int next2 = next + (int)gridDim.x;
        if (next2 < n_tiles) { const int64_t r2 = (int64_t)next2 * TR; bo0 = kc.offsets[r2]; bo1 = kc.offsets[r2 + tile_rows(next2)]; }
      }
    }
    if (next < n_tiles) prefetch(next);
    // ---- aggregate arguments of THIS tile: in flight while the keys are hashed and probed ----
    unsigned long long av[HS_PRE][R];
#pragma unroll
    for (int i = 0; i < HS_PRE; ++i) {
      if (pre_acc[i] >= 0) {
        const unsigned long long* as = (const unsigned long long*)P.cols[P.accs[pre_acc[i]].arg_slot].data + row0 + lr0;
        ld_vec_u64<R>(as, full && (reinterpret_cast<uintptr_t>(as) & 15) == 0, lr0, rows, av[i]);
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) av[i][j] = 0;
      }
    }
    // ---- predicate ----
    unsigned ok = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      bool f = lr0 + j < rows;
      if (PRED == 1 && f) {
        const ColView& c = P.cols[P.sp_slot];
        f = cmp_i64(P.sp_cmp, P.sp_is_f64 ? f64_total_key(pv[j]) : (long long)pv[j], pred_c) && col_valid(c, row0 + lr0 + j);
      }
      ok |= (unsigned)f << j;
    }
    // ---- keys ----
    Key16 mine[R];
    unsigned home[R];
    if (KEYK == KEY_BYTES) {
      off[R] = __shfl_down_sync(0xffffffffu, off[0], 1);
      if ((lane == 31 || lr0 + R >= rows) && lr0 + R <= rows) off[R] = offx;
      const bool staged = s_str_staged[st];
      const int base = s_str_base[st];
      if (staged) { mbar_wait(&s_bar[st], (ph >> st) & 1); ph ^= 1u << st; }
      const uint8_t* in_bytes = smem + st * stage_bytes;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        home[j] = 0;
        if (!((ok >> j) & 1)) continue;
        const int64_t row = row0 + lr0 + j;
        unsigned h32;
        if (!col_valid(kc, row)) { mine[j].lo = 0; mine[j].hi = (unsigned long long)KEYTAG_NULL << 32; h32 = hash32_key16(mine[j]); }
        else if (staged) { make_key_smem(in_bytes + (off[j] - base), off[j + 1] - off[j], row, &mine[j], &h32); if (off[j + 1] - off[j] > 12) long_flag = 1; }
        else {
          int llen = 0;
          const uint8_t* lp = make_key_raw(KEY_BYTES, kc, row, &mine[j], &llen);
          if (lp) { const unsigned long long h = hash_bytes(lp, llen); h32 = (unsigned)(h >> 32) ^ (unsigned)h; long_flag = 1; }
          else h32 = hash32_key16(mine[j]);
        }
        home[j] = (unsigned)((h32 * 0x9E3779B1u) & bmask);
      }
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        home[j] = 0;
        if (!((ok >> j) & 1)) continue;
        if (col_valid(kc, row0 + lr0 + j)) { mine[j].lo = kv[j]; mine[j].hi = (unsigned long long)KEYTAG_INT << 32; }
        else { mine[j].lo = 0; mine[j].hi = (unsigned long long)KEYTAG_NULL << 32; }
        home[j] = (unsigned)((hash32_key16(mine[j]) * 0x9E3779B1u) & bmask);
      }
    }
    // this stage's key bytes are in registers: the producer may refill it in the iteration after next.  The barrier
    // also carries the producer's poll of the overflow flag to every thread (CTA-uniform exit).
    if (__syncthreads_or(stop)) {
      // a bulk copy for the next tile may be in flight into this CTA's shared memory: let it land before leaving
      if (KEYK == KEY_BYTES && next < n_tiles && s_str_staged[st ^ 1]) mbar_wait(&s_bar[st ^ 1], (ph >> (st ^ 1)) & 1);
      break;
    }
    // ---- home buckets of all R rows requested at once ----
    Key16 kb[R][TBL_B];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!((ok >> j) & 1)) continue;
      if (dbg & 4) { kb[j][0].lo = home[j]; continue; }  // measurement: no table reads at all
      const Key16* b = reinterpret_cast<const Key16*>(P.table + (unsigned long long)home[j] * (unsigned long long)bstride);
      ld256_keys(b, &kb[j][0], &kb[j][1]);
      ld256_keys(b + 2, &kb[j][2], &kb[j][3]);
    }
    unsigned long long slot[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      slot[j] = ~0ull;
      if (!((ok >> j) & 1)) continue;
      if (dbg & 2) { slot[j] = (unsigned long long)home[j] * TBL_B + (kb[j][0].lo & 3); continue; }  // measurement: no probing
      Key16* b = reinterpret_cast<Key16*>(P.table + (unsigned long long)home[j] * (unsigned long long)bstride);
      bool done = false;
#pragma unroll
      for (int i = 0; i < TBL_B; ++i) {
        if (done) continue;
        Key16 c = kb[j][i];
        if (c.hi == KEY_EMPTY) {
          c = cas128(b + i, Key16{KEY_EMPTY, KEY_EMPTY}, mine[j]);
          if (c.hi == KEY_EMPTY && c.lo == KEY_EMPTY) { ++claimed; done = true; slot[j] = (unsigned long long)home[j] * TBL_B + i; continue; }
        }
        if (key_equal(mine[j], c, kc, kc)) { done = true; slot[j] = (unsigned long long)home[j] * TBL_B + i; }
      }
      if (!done) {  // the home bucket is full of other keys: walk on
        slot[j] = table_find_or_claim(P.table, bmask, bstride, (unsigned long long)home[j] + 1, mine[j], kc, kc, &claimed);
        if (slot[j] == ~0ull) { err_overflow = 1; ok &= ~(1u << j); }  // table too loaded for this batch
      }
    }
    // ---- accumulate (fire-and-forget REDs) ----
    if (dbg & 1) continue;  // measurement: no accumulation
    if (SIG == 1) {
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (!((ok >> j) & 1)) continue;
        atomicAdd(tbl_acc(P.table, slot[j], 0, bstride), 1ull);
        atomicAdd(tbl_acc(P.table, slot[j], 1, bstride), av[0][j]);
      }
    } else {
      for (int a = 0; a < P.n_acc; ++a) {
        const AccParam& A = P.accs[a];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (!((ok >> j) & 1)) continue;
          unsigned long long bits = 0;
          bool valid = true;
          if (A.kind != ACC_COUNT_STAR) {
            const ColView& c = P.cols[A.arg_slot];
            valid = col_valid(c, row0 + lr0 + j);
            bits = a == pre_acc[0] ? av[0][j] : (a == pre_acc[1] ? av[1][j] : __ldcs((const unsigned long long*)c.data + row0 + lr0 + j));
          }
          if (valid) accumulate(A.kind, A.arg_is_f64, tbl_acc(P.table, slot[j], a, bstride), bits);
        }
      }
    }
  }
  if (err_overflow) atomicExch(P.overflow, 1);
  if (long_flag) *P.long_seen = 1;
  claimed = (unsigned int)__reduce_add_sync(0xffffffffu, claimed);
  if (lane == 0 && claimed) atomicAdd(P.group_count, claimed);
}

}  // namespace

// Returns false when the plan / batch shape is not covered (the caller then uses hash_agg_kernel).
bool launch_hash_agg_stream(const AggParams& P, unsigned long long capacity, int64_t key_bytes, cudaStream_t stream) {
  static const int enabled = [] { const char* e = getenv("ARK_AGG_STREAM"); return e ? atoi(e) : 1; }();
  if (!enabled) return false;
  if (P.pred_kind == 2 || (P.key_kind != KEY_BYTES && P.key_kind != KEY_INT64)) return false;
  if (capacity > (1ull << 31) || P.n_rows <= 0) return false;
  for (int a = 0; a < P.n_acc; ++a) if (P.accs[a].arg_prog >= 0) return false;
  if (P.pred_kind == 1 && (reinterpret_cast<uintptr_t>(P.cols[P.sp_slot].data) & 7)) return false;
  static const int rows_per_thread = [] { const char* e = getenv("ARK_AGG_STREAM_R"); const int v = e ? atoi(e) : 2; return v == 1 || v == 4 ? v : 2; }();
  const int R = rows_per_thread;
  const int TR = HS_THREADS * R;
  int cap = 0;
  if (P.key_kind == KEY_BYTES) {
    if (key_bytes < 0) return false;  // the caller resolves the key column's extent first
    const double avg = (double)key_bytes / (double)P.n_rows;
    cap = (int)round_up((int64_t)(avg * TR * 1.0625) + 64, 1024);
    cap = std::max(2048, std::min(cap, 48 * 1024));
  }
  const size_t smem = cap ? 2 * (size_t)(cap + 32) : 0;
  // {COUNT(*), SUM(non-null Int64 column)}: the specialised accumulate
  const bool sig1 = P.n_acc == 2 && P.accs[0].kind == ACC_COUNT_STAR && P.accs[1].kind == ACC_SUM_I64 && P.cols[P.accs[1].arg_slot].validity == nullptr;
  const void* fn = nullptr;
#define ARK_HS_R(K, PR, S) (R == 1 ? (const void*)hash_agg_stream_kernel<K, PR, 1, S> : R == 4 ? (const void*)hash_agg_stream_kernel<K, PR, 4, S> : (const void*)hash_agg_stream_kernel<K, PR, 2, S>)
#define ARK_HS_FN(K, PR) (sig1 ? ARK_HS_R(K, PR, 1) : ARK_HS_R(K, PR, 0))
  if (P.key_kind == KEY_BYTES) fn = P.pred_kind ? ARK_HS_FN(KEY_BYTES, 1) : ARK_HS_FN(KEY_BYTES, 0);
  else fn = P.pred_kind ? ARK_HS_FN(KEY_INT64, 1) : ARK_HS_FN(KEY_INT64, 0);
#undef ARK_HS_FN
#undef ARK_HS_R
  ARK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (48 * 1024 + 32)));
  int occ = 0;
  ARK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, HS_THREADS, smem));
  if (occ < 1) return false;
  static const int sms = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
  static const int cap_per_sm = [] { const char* e = getenv("ARK_AGG_STREAM_CTAS"); return e ? atoi(e) : 0; }();
  if (cap_per_sm > 0) occ = std::min(occ, cap_per_sm);
  const int n_tiles = (int)ceil_div(P.n_rows, TR);
  const int grid = std::max(1, std::min(n_tiles, sms * occ));
  KernelTimer t("hash_agg_kernel", stream);
  static const int dbg = [] { const char* e = getenv("ARK_AGG_DEBUG"); return e ? atoi(e) : 0; }();  // measurement knob, results void
  void* args[] = {(void*)&P, (void*)&cap, (void*)&dbg};
  ARK_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(HS_THREADS), args, smem, stream));
  return true;
}

}  // namespace ark
