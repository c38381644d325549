// hash_agg_tile.cu — tiled GROUP BY kernel: the common shape (optional `col <cmp> literal` filter,
// column-valued aggregate arguments) restructured for memory-level parallelism.
//
// hash_agg_kernel (hash_agg.cu) walks one row per thread through a chain of dependent loads
// (offsets → key bytes → table slot → atomics); ncu shows it latency-bound (issue slots 20 % busy,
// long-scoreboard stalls 39 per issue).  Here a CTA takes 1024-row tiles and each thread owns 4
// consecutive rows:
//   * the tile's key bytes arrive through ONE 1-D TMA bulk copy (cp.async.bulk → mbarrier) of the
//     16-byte-aligned window around [offsets[t0], offsets[t1]), so keys are built from shared memory;
//   * offsets / predicate / argument values are 16-byte vector loads issued up front;
//   * the 4 table slots of a thread are fetched with 4 independent 128-bit loads before any is resolved;
//   * low-cardinality tables (≤ 2048 slots) accumulate in a per-CTA shared-memory copy of the
//     accumulators (shared-memory atomics) that is flushed once per CTA — hot keys no longer
//     serialise on L2 atomics (K = 2: 12.3 ms → see profiles/).
// Table layout, key format and CAS protocol are those of hash_agg.cu.
#include <atomic>

#include "agg_acc.cuh"
#include "engine.h"
#include "hash_agg.cuh"
#include "hashkey.cuh"

namespace ark {

namespace {

constexpr int HT_THREADS = 256;
constexpr int HT_TILE = HT_THREADS * 4;

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
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

// Key16 + hash of a byte string that sits in shared memory (same encoding as make_key)
__device__ __forceinline__ void make_key_smem(const uint8_t* p, int len, int64_t row, Key16* key, unsigned long long* hash) {
  Key16 k;
  if (len <= 12) {
    unsigned w[3] = {0, 0, 0};
    if ((smem_addr(p) & 3) == 0) {
      const unsigned* q = reinterpret_cast<const unsigned*>(p);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int rem = len - 4 * i;
        if (rem >= 4) w[i] = q[i];
        else if (rem > 0) { for (int b = 0; b < rem; ++b) w[i] |= (unsigned)p[4 * i + b] << (8 * b); }
      }
    } else {
      for (int b = 0; b < len; ++b) w[b >> 2] |= (unsigned)p[b] << (8 * (b & 3));
    }
    k.lo = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
    k.hi = (unsigned long long)w[2] | ((unsigned long long)(unsigned)len << 32);
    *key = k; *hash = hash_key16(k);
  } else {
    const unsigned prefix = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
    k.lo = (unsigned long long)row;
    k.hi = (unsigned long long)prefix | ((unsigned long long)(KEYTAG_LONG | (unsigned)len) << 32);
    *key = k; *hash = hash_bytes(p, len);
  }
}

template <int PRED>
__global__ void __launch_bounds__(HT_THREADS, 4) hash_agg_tile_kernel(const __grid_constant__ AggParams P, const int str_cap, const int priv_slots) {
  extern __shared__ __align__(16) uint8_t smem[];  // [key bytes window: str_cap + 32][privatised accumulators]
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_str_base, s_str_staged, s_stop;
  const int tid = threadIdx.x, lane = tid & 31;
  uint8_t* in_bytes = smem;
  unsigned long long* priv = reinterpret_cast<unsigned long long*>(smem + (str_cap ? str_cap + 32 : 0));
  const int64_t n = P.n_rows;
  const int n_tiles = (int)((n + HT_TILE - 1) / HT_TILE);
  const bool bytes_key = P.key_kind == KEY_BYTES;
  const ColView& kc = P.cols[P.key_kind == KEY_NONE ? 0 : P.key_slot];
  const int32_t* koff = kc.offsets;

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (P.key_kind == KEY_NONE && blockIdx.x == 0) {  // a global aggregate always yields one row
      Key16 mine; unsigned long long h;
      make_key(KEY_NONE, kc, 0, &mine, &h);
      Key16 cur = cas128(reinterpret_cast<Key16*>(P.table + (h & P.mask) * (unsigned long long)P.slot_stride), Key16{KEY_EMPTY, KEY_EMPTY}, mine);
      if (cur.hi == KEY_EMPTY && cur.lo == KEY_EMPTY) atomicAdd(P.group_count, 1u);
    }
  }
  for (int i = tid; i < priv_slots * P.n_acc; i += HT_THREADS) priv[i] = acc_identity(P.accs[i % P.n_acc].kind);
  __syncthreads();

  const long long pred_c = P.sp_is_f64 ? f64_total_key(P.sp_const) : (long long)P.sp_const;
  for (int tile = blockIdx.x, it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int64_t row0 = (int64_t)tile * HT_TILE;
    const int rows = (int)((n - row0) < HT_TILE ? (n - row0) : HT_TILE);
    if (tid == 0) s_stop = *reinterpret_cast<volatile int32_t*>(P.overflow);  // table too small: the host retries with 4× the slots
    if (bytes_key && tid == 0 && !s_stop) {
      const int32_t o0 = koff[row0], o1 = koff[row0 + rows];
      const uintptr_t a0 = reinterpret_cast<uintptr_t>((const uint8_t*)kc.data + o0), a1 = reinterpret_cast<uintptr_t>((const uint8_t*)kc.data + o1);
      const uintptr_t lo = a0 & ~(uintptr_t)15, hi = (a1 + 15) & ~(uintptr_t)15;
      int staged = 0;
      if (o1 > o0 && hi - lo <= (uintptr_t)str_cap) {
        staged = 1;
        mbar_expect_tx(&s_bar, (unsigned)(hi - lo));
        tma_load_1d(in_bytes, reinterpret_cast<const void*>(lo), (unsigned)(hi - lo), &s_bar);
      }
      s_str_base = o0 - (int32_t)(a0 - lo); s_str_staged = staged;
    }
    // ---- loads in flight: offsets, predicate column ----
    const int lr0 = 4 * tid;
    const bool full = lr0 + 4 <= rows;
    int off[5] = {0, 0, 0, 0, 0};
    unsigned long long pv[4] = {0, 0, 0, 0};
    if (bytes_key) {
      const int32_t* os = koff + row0 + lr0;
      if (full && (reinterpret_cast<uintptr_t>(os) & 15) == 0) {
        asm volatile("ld.global.cs.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(off[0]), "=r"(off[1]), "=r"(off[2]), "=r"(off[3]) : "l"(os));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) off[j] = (lr0 + j <= rows) ? os[j] : 0;
      }
      off[4] = __shfl_down_sync(0xffffffffu, off[0], 1);
      if ((lane == 31 || lr0 + 4 >= rows) && lr0 + 4 <= rows) off[4] = os[4];
    }
    if (PRED == 1) {
      const unsigned long long* src = (const unsigned long long*)P.cols[P.sp_slot].data + row0 + lr0;
      if (full && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(pv[0]), "=l"(pv[1]) : "l"(src));
        asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(pv[2]), "=l"(pv[3]) : "l"(src + 2));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) pv[j] = (lr0 + j < rows) ? src[j] : 0;
      }
    }
    __syncthreads();  // s_str_staged / s_str_base / s_stop visible; (it > 0) previous tile's readers are done
    if (s_stop) break;  // CTA-uniform; no bulk copy was issued for this tile
    unsigned ok = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool f = lr0 + j < rows;
      if (PRED == 1 && f) {
        const ColView& c = P.cols[P.sp_slot];
        f = cmp_i64(P.sp_cmp, P.sp_is_f64 ? f64_total_key(pv[j]) : (long long)pv[j], pred_c) && col_valid(c, row0 + lr0 + j);
      }
      ok |= (unsigned)f << j;
    }
    const bool staged = bytes_key && s_str_staged;
    if (staged) mbar_wait(&s_bar, it & 1);
    // ---- keys + hashes, then all 4 table slots fetched before any is resolved ----
    Key16 mine[4];
    unsigned long long slot[4];
    Key16 cur[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      slot[j] = 0;
      if (!((ok >> j) & 1)) continue;
      const int64_t row = row0 + lr0 + j;
      unsigned long long h;
      if (bytes_key && staged && col_valid(kc, row)) make_key_smem(in_bytes + (off[j] - s_str_base), off[j + 1] - off[j], row, &mine[j], &h);
      else make_key(P.key_kind, kc, row, &mine[j], &h);
      slot[j] = h & P.mask;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((ok >> j) & 1) cur[j] = ld128(reinterpret_cast<const Key16*>(P.table + slot[j] * (unsigned long long)P.slot_stride));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!((ok >> j) & 1)) continue;
      Key16 c = cur[j];
      int probes = 0;
      while (true) {
        Key16* sk = reinterpret_cast<Key16*>(P.table + slot[j] * (unsigned long long)P.slot_stride);
        if (c.hi == KEY_EMPTY) {
          c = cas128(sk, Key16{KEY_EMPTY, KEY_EMPTY}, mine[j]);
          if (c.hi == KEY_EMPTY && c.lo == KEY_EMPTY) {
            const unsigned g = atomicAdd(P.group_count, 1u);
            if (g >= P.max_groups) atomicExch(P.overflow, 1);
            break;
          }
        }
        if (key_equal(mine[j], c, kc, kc)) break;
        slot[j] = (slot[j] + 1) & P.mask;
        if (++probes > 4096) { atomicExch(P.overflow, 1); ok &= ~(1u << j); break; }
        c = ld128(reinterpret_cast<const Key16*>(P.table + slot[j] * (unsigned long long)P.slot_stride));
      }
    }
    // ---- accumulate ----
    for (int a = 0; a < P.n_acc; ++a) {
      const AccParam& A = P.accs[a];
      unsigned long long av[4] = {0, 0, 0, 0};
      unsigned valid = ok;
      if (A.kind != ACC_COUNT_STAR) {
        const ColView& c = P.cols[A.arg_slot];
        const unsigned long long* src = (const unsigned long long*)c.data + row0 + lr0;
        if (full && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(av[0]), "=l"(av[1]) : "l"(src));
          asm volatile("ld.global.cs.v2.u64 {%0, %1}, [%2];" : "=l"(av[2]), "=l"(av[3]) : "l"(src + 2));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) av[j] = (lr0 + j < rows) ? src[j] : 0;
        }
        if (c.validity) {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (((valid >> j) & 1) && !col_valid(c, row0 + lr0 + j)) valid &= ~(1u << j);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!((valid >> j) & 1)) continue;
        unsigned long long* dst = priv_slots ? priv + slot[j] * P.n_acc + a
                                             : reinterpret_cast<unsigned long long*>(P.table + slot[j] * (unsigned long long)P.slot_stride + A.acc_offset);
        accumulate(A.kind, A.arg_is_f64, dst, av[j]);
      }
    }
    __syncthreads();  // everyone is done with in_bytes before the next tile's bulk copy overwrites it
  }
  // ---- flush the privatised accumulators ----
  if (priv_slots) {
    for (int i = tid; i < priv_slots * P.n_acc; i += HT_THREADS) {
      const int a = i % P.n_acc, s = i / P.n_acc;
      const unsigned long long v = priv[i];
      if (v != acc_identity(P.accs[a].kind))
        merge_acc(P.accs[a].kind, reinterpret_cast<unsigned long long*>(P.table + (unsigned long long)s * P.slot_stride + P.accs[a].acc_offset), v);
    }
  }
}

}  // namespace

// Returns false when the plan/batch shape is not covered (caller uses hash_agg_kernel).
bool launch_hash_agg_tile(const AggParams& P, unsigned long long capacity, int64_t key_bytes, cudaStream_t stream) {
  if (P.pred_kind == 2) return false;
  for (int a = 0; a < P.n_acc; ++a) if (P.accs[a].arg_prog >= 0) return false;
  const int64_t n = P.n_rows;
  int cap = 0;
  if (P.key_kind == KEY_BYTES) {
    static std::atomic<double> avg_hint{12.8};
    const double avg = (n > 0 && key_bytes >= 0) ? (double)key_bytes / (double)n : avg_hint.load();
    if (n > 0 && key_bytes >= 0) avg_hint.store(avg);
    cap = (int)round_up((int64_t)(avg * HT_TILE * 1.0625) + 64, 1024);
    cap = std::max(4096, std::min(cap, 48 * 1024));
  }
  int priv_slots = 0;
  if (capacity <= 2048 && P.n_acc <= 6) priv_slots = (int)capacity;
  const size_t smem = (cap ? cap + 32 : 0) + (size_t)priv_slots * P.n_acc * 8;
  static bool configured = false;
  if (!configured) {
    ARK_CUDA(cudaFuncSetAttribute(hash_agg_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    ARK_CUDA(cudaFuncSetAttribute(hash_agg_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  const int n_tiles = (int)ceil_div(n, HT_TILE);
  const int grid = std::max(1, std::min(n_tiles, 148 * 4));
  KernelTimer t("hash_agg_tile_kernel", stream);
  if (P.pred_kind == 0) hash_agg_tile_kernel<0><<<grid, HT_THREADS, smem, stream>>>(P, cap, priv_slots);
  else hash_agg_tile_kernel<1><<<grid, HT_THREADS, smem, stream>>>(P, cap, priv_slots);
  return true;
}

}  // namespace ark
