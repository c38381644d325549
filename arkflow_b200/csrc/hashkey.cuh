// hashkey.cuh — 16-byte table keys shared by the GROUP BY and JOIN kernels: construction from a
// column row, hashing, equality, and the 128-bit CAS / load used to claim and read table slots.
#pragma once
#include "hash_agg.cuh"
#include "tma.cuh"
#include "vm.cuh"

namespace ark {

static __device__ __forceinline__ Key16 cas128(Key16* addr, Key16 cmp, Key16 val) {
  Key16 old;
  asm volatile(
      "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
      "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr)
      : "memory");
  return old;
}
static __device__ __forceinline__ Key16 ld128(const Key16* addr) {  // single 128-bit access (LDG.E.128.STRONG.GPU)
  Key16 v;
  asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}"
               : "=l"(v.lo), "=l"(v.hi) : "l"(addr) : "memory");
  return v;
}

// two adjacent keys (32 bytes, 32-byte aligned) with one 256-bit access (LDG.E.256.STRONG.GPU): one L2 request per sector
static __device__ __forceinline__ void ld256_keys(const Key16* addr, Key16* k0, Key16* k1) {
  asm volatile("ld.relaxed.gpu.global.v4.u64 {%0, %1, %2, %3}, [%4];"
               : "=l"(k0->lo), "=l"(k0->hi), "=l"(k1->lo), "=l"(k1->hi) : "l"(addr) : "memory");
}

static __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static __device__ __forceinline__ unsigned long long hash_key16(Key16 k) {
  return mix64(k.lo * 0x9E3779B97F4A7C15ull + mix64(k.hi + 0x632BE59BD9B4E019ull));
}
static __device__ inline unsigned long long hash_bytes(const uint8_t* p, int len) {
  unsigned long long h = 0xCBF29CE484222325ull;
  for (int i = 0; i < len; ++i) { h ^= p[i]; h *= 0x100000001B3ull; }
  return mix64(h ^ (unsigned long long)len);
}

// Builds the table key of (column c, row).  Returns the key's bytes (and *long_len) when it is a long,
// out-of-line key, nullptr otherwise.
static __device__ inline const uint8_t* make_key_raw(int kind, const ColView& c, int64_t row, Key16* key, int* long_len) {
  Key16 k;
  if (kind == KEY_NONE) { k.lo = 0; k.hi = (unsigned long long)KEYTAG_INT << 32; *key = k; return nullptr; }
  if (!col_valid(c, row)) { k.lo = 0; k.hi = (unsigned long long)KEYTAG_NULL << 32; *key = k; return nullptr; }
  if (kind == KEY_INT64) {
    k.lo = __ldcs((const unsigned long long*)c.data + row); k.hi = (unsigned long long)KEYTAG_INT << 32;  // streaming: keep L2 for the table
  } else if (kind == KEY_BOOL) {
    k.lo = bit_get((const uint8_t*)c.data, row + c.data_bit0); k.hi = (unsigned long long)KEYTAG_INT << 32;
  } else {
    const int32_t o0 = __ldcs(c.offsets + row), o1 = __ldcs(c.offsets + row + 1);
    const int len = o1 - o0;
    const uint8_t* p = (const uint8_t*)c.data + o0;
    if (len <= 12) {
      unsigned w[3] = {0, 0, 0};
      if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
        const unsigned* q = (const unsigned*)p;
        for (int i = 0; i < 3; ++i) {
          const int rem = len - 4 * i;
          if (rem >= 4) w[i] = __ldcs(q + i);
          else if (rem > 0) { for (int b = 0; b < rem; ++b) w[i] |= (unsigned)p[4 * i + b] << (8 * b); }
        }
      } else {
        for (int b = 0; b < len; ++b) w[b >> 2] |= (unsigned)p[b] << (8 * (b & 3));
      }
      k.lo = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
      k.hi = (unsigned long long)w[2] | ((unsigned long long)(unsigned)len << 32);
    } else {
      unsigned prefix = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
      k.lo = (unsigned long long)row;
      k.hi = (unsigned long long)prefix | ((unsigned long long)(KEYTAG_LONG | (unsigned)len) << 32);
      *key = k; *long_len = len; return p;
    }
  }
  *key = k;
  return nullptr;
}

// Key16 + 64-bit hash.
static __device__ inline void make_key(int kind, const ColView& c, int64_t row, Key16* key, unsigned long long* hash) {
  int len = 0;
  const uint8_t* p = make_key_raw(kind, c, row, key, &len);
  *hash = kind == KEY_NONE ? 0 : (p ? hash_bytes(p, len) : hash_key16(*key));
}

// 32-bit hash of a short key (~16 instructions; hash_key16 costs ~60): the partitioned GROUP BY takes its
// bucket from the top bits and the slot inside the bucket's region from the low bits.
static __device__ __forceinline__ unsigned hash32_key16(Key16 k) {
  unsigned h = (unsigned)k.lo * 0x85EBCA6Bu;
  h ^= __funnelshift_l((unsigned)(k.lo >> 32) * 0xC2B2AE35u, (unsigned)(k.lo >> 32) * 0xC2B2AE35u, 13);
  h ^= __funnelshift_l((unsigned)k.hi * 0x27D4EB2Fu, (unsigned)k.hi * 0x27D4EB2Fu, 7);
  h ^= (unsigned)(k.hi >> 32) * 0x165667B1u;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

static __device__ __forceinline__ bool key_is_long(Key16 k) {
  const unsigned tag = (unsigned)(k.hi >> 32);
  return (tag & KEYTAG_LONG) && tag < KEYTAG_NULL;
}

// equality of a probing key with a stored key (long keys: compare the bytes of both rows)
// `mine` was built from column mc, `stored` from column sc (the same column for GROUP BY; probe vs
// build column for a join)
static __device__ inline bool key_equal(Key16 mine, Key16 stored, const ColView& mc, const ColView& sc) {
  if (!key_is_long(mine)) return mine.lo == stored.lo && mine.hi == stored.hi;
  if (mine.hi != stored.hi) return false;
  if (mine.lo == stored.lo && mc.data == sc.data && mc.offsets == sc.offsets) return true;
  const int len = (int)((unsigned)(mine.hi >> 32) & 0x7FFFFFFFu);
  const uint8_t* a = (const uint8_t*)mc.data + mc.offsets[(int64_t)mine.lo];
  const uint8_t* b = (const uint8_t*)sc.data + sc.offsets[(int64_t)stored.lo];
  for (int i = 0; i < len; ++i) if (a[i] != b[i]) return false;
  return true;
}


static __device__ __forceinline__ bool key_is_pair(Key16 k) { return (unsigned)(k.hi >> 32) == KEYTAG_PAIR; }

// equality of the values two rows hold in one key column (NULL equals NULL: one group)
static __device__ inline bool rows_equal(int kind, const ColView& c, int64_t a, int64_t b) {
  const bool va = col_valid(c, a), vb = col_valid(c, b);
  if (!va || !vb) return va == vb;
  if (kind == KEY_INT64) return ((const unsigned long long*)c.data)[a] == ((const unsigned long long*)c.data)[b];
  if (kind == KEY_BOOL) return bit_get((const uint8_t*)c.data, a + c.data_bit0) == bit_get((const uint8_t*)c.data, b + c.data_bit0);
  const int32_t a0 = c.offsets[a], b0 = c.offsets[b];
  const int la = c.offsets[a + 1] - a0, lb = c.offsets[b + 1] - b0;
  if (la != lb) return false;
  const uint8_t* pa = (const uint8_t*)c.data + a0;
  const uint8_t* pb = (const uint8_t*)c.data + b0;
  for (int i = 0; i < la; ++i) if (pa[i] != pb[i]) return false;
  return true;
}

static __device__ __forceinline__ unsigned long long stored_key_hash(Key16 k, const ColView& kc) {
  if (key_is_pair(k)) return (k.hi & 0xFFFFFFFFull) << 32;  // the stored top half of the pair's hash: enough for partition_of
  if (key_is_long(k)) {
    const int len = (int)((unsigned)(k.hi >> 32) & 0x7FFFFFFFu);
    return hash_bytes((const uint8_t*)kc.data + kc.offsets[(int64_t)k.lo], len);
  }
  return hash_key16(k);
}
static __device__ __forceinline__ unsigned stored_key_hash32(Key16 k, const ColView& kc) {
  if (key_is_long(k)) {
    const int len = (int)((unsigned)(k.hi >> 32) & 0x7FFFFFFFu);
    const unsigned long long h = hash_bytes((const uint8_t*)kc.data + kc.offsets[(int64_t)k.lo], len);
    return (unsigned)(h >> 32) ^ (unsigned)h;
  }
  return hash32_key16(k);
}
// owner partition of a key hash: top 24 bits scaled to [0, n_parts) — identical on every rank
static __device__ __forceinline__ int partition_of(unsigned long long h, int n_parts) {
  return (int)(((h >> 40) * (unsigned long long)n_parts) >> 24);
}

// Key16 + 32-bit table hash of a byte string that sits in shared memory (same key encoding as make_key)
static __device__ __forceinline__ void make_key_smem(const uint8_t* p, int len, int64_t row, Key16* key, unsigned int* hash) {
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
    *key = k; *hash = hash32_key16(k);
  } else {
    const unsigned prefix = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
    k.lo = (unsigned long long)row;
    k.hi = (unsigned long long)prefix | ((unsigned long long)(KEYTAG_LONG | (unsigned)len) << 32);
    *key = k;
    const unsigned long long h = hash_bytes(p, len);
    *hash = (unsigned)(h >> 32) ^ (unsigned)h;
  }
}


// The table key of a (k1, k2) pair at `row` and its 64-bit hash.
static __device__ inline unsigned long long make_pair_key(int kind1, const ColView& c1, int kind2, const ColView& c2, int64_t row, Key16* key) {
  Key16 t; unsigned long long h1, h2;
  make_key(kind1, c1, row, &t, &h1);
  make_key(kind2, c2, row, &t, &h2);
  const unsigned long long h = mix64(h1 * 0x9E3779B97F4A7C15ull + (h2 ^ 0xD6E8FEB86659FD93ull));
  key->lo = (unsigned long long)row;
  key->hi = (h >> 32) | ((unsigned long long)KEYTAG_PAIR << 32);
  return h;
}

// table_find_or_claim for pair keys: equal fingerprints, then the two columns of the two rows are compared
static __device__ __forceinline__ unsigned long long table_find_or_claim_pair(uint8_t* table, unsigned long long bucket_mask, int bstride,
                                                                              unsigned long long home, Key16 mine, int kind1, const ColView& c1,
                                                                              int kind2, const ColView& c2, unsigned int* claimed, int max_buckets = 128) {
  unsigned long long b = home & bucket_mask;
  for (int probes = 0; probes < max_buckets; ++probes) {
    Key16* kb = reinterpret_cast<Key16*>(table + b * (unsigned long long)bstride);
    Key16 k[TBL_B];
    ld256_keys(kb, &k[0], &k[1]);
    ld256_keys(kb + 2, &k[2], &k[3]);
#pragma unroll
    for (int i = 0; i < TBL_B; ++i) {
      Key16 c = k[i];
      if (c.hi == KEY_EMPTY) {
        c = cas128(kb + i, Key16{KEY_EMPTY, KEY_EMPTY}, mine);
        if (c.hi == KEY_EMPTY && c.lo == KEY_EMPTY) { ++*claimed; return b * TBL_B + i; }
      }
      if (c.hi == mine.hi && (c.lo == mine.lo || (rows_equal(kind1, c1, (int64_t)mine.lo, (int64_t)c.lo) && rows_equal(kind2, c2, (int64_t)mine.lo, (int64_t)c.lo))))
        return b * TBL_B + i;
    }
    b = (b + 1) & bucket_mask;
  }
  return ~0ull;
}

// Finds the slot of `mine` in the bucketed table or claims the first free lane on its probe path (buckets home,
// home + 1, …; lanes 0..3 in order, so a key is always met before any empty lane after it).  `mc` / `sc`: the
// columns long keys of the prober / of stored keys point into.  Returns the slot, or ~0 when `max_buckets` buckets
// were full of other keys (table too loaded: the host retries with more slots).
static __device__ __forceinline__ unsigned long long table_find_or_claim(uint8_t* table, unsigned long long bucket_mask, int bstride,
                                                                         unsigned long long home, Key16 mine, const ColView& mc, const ColView& sc,
                                                                         unsigned int* claimed, int max_buckets = 128) {
  unsigned long long b = home & bucket_mask;
  for (int probes = 0; probes < max_buckets; ++probes) {
    Key16* kb = reinterpret_cast<Key16*>(table + b * (unsigned long long)bstride);
    Key16 k[TBL_B];
    ld256_keys(kb, &k[0], &k[1]);
    ld256_keys(kb + 2, &k[2], &k[3]);
#pragma unroll
    for (int i = 0; i < TBL_B; ++i) {
      Key16 c = k[i];
      if (c.hi == KEY_EMPTY) {
        c = cas128(kb + i, Key16{KEY_EMPTY, KEY_EMPTY}, mine);
        if (c.hi == KEY_EMPTY && c.lo == KEY_EMPTY) { ++*claimed; return b * TBL_B + i; }
      }
      if (key_equal(mine, c, mc, sc)) return b * TBL_B + i;
    }
    b = (b + 1) & bucket_mask;
  }
  return ~0ull;
}

}  // namespace ark
