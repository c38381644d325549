// filter_project.cuh — parameter block of the fused filter + project + compaction kernel.
#pragma once
#include "vm.h"

namespace ark {

constexpr int FP_THREADS = 256;
constexpr int FP_CHUNKS = 4;                              // chunks per tile, 2 rows per thread per chunk
constexpr int FP_CHUNK_ROWS = FP_THREADS * 2;             // 512
constexpr int FP_TILE = FP_CHUNK_ROWS * FP_CHUNKS;        // 2048 rows per CTA tile
constexpr int FP_WARPS = FP_THREADS / 32;
constexpr int FP_MAX_OUT = 12;
constexpr int FP_MAX_PROGS = 3;
constexpr int FP_MAX_VARLEN = 2;                          // var-len outputs per launch
constexpr int FP_CHANNELS = 1 + FP_MAX_VARLEN;            // scan channels: rows, bytes of var-len 0/1
constexpr int FP_STR_STAGE = 36 * 1024;                   // bytes of shared memory staging per tile

enum FpOutKind : int32_t {
  FP_OUT_FIXED8 = 0,    // 8-byte passthrough of cols[slot]
  FP_OUT_VARLEN = 1,    // Utf8/Binary passthrough of cols[slot]; scan channel 1 + varlen_idx
  FP_OUT_COMPUTED8 = 2, // progs[prog] result, 8 bytes
  FP_OUT_BOOL = 3,      // Boolean passthrough → byte per row (packed by pack_bits_kernel)
  FP_OUT_COMPUTED_BOOL = 4,
};

struct FpOutput {
  int32_t kind;
  int32_t slot;
  int32_t prog;
  int32_t varlen_idx;
  int32_t write_validity;  // 1 ⇒ out_valid receives one byte per output row
  int32_t pad;
  void* out_data;          // 8-byte values | string bytes | byte-per-row booleans
  int32_t* out_offsets;    // var-len: n_out+1 offsets (the kernel writes [0, total]; host patches none)
  uint8_t* out_valid;      // byte-per-row validity
};

struct FpParams {
  int64_t n_rows;
  int32_t n_tiles;
  int32_t n_out;
  int32_t n_varlen;
  int32_t varlen_slot[FP_MAX_VARLEN];  // column slot of var-len output v
  int32_t pad0;
  // predicate: 0 = none, 1 = simple (col cmp literal), 2 = VmProgram
  int32_t sp_slot, sp_cmp, sp_is_f64;
  uint64_t sp_const;
  ColView cols[MAX_COLS];
  FpOutput outs[FP_MAX_OUT];
  VmProgram pred;
  VmProgram progs[FP_MAX_PROGS];
  unsigned long long* desc;  // [n_tiles][FP_CHANNELS] look-back descriptors (zeroed before launch)
  unsigned int* ticket;      // dynamic tile counter (zeroed before launch)
  long long* totals;         // [FP_CHANNELS] inclusive totals written by the last tile
  int32_t* error;            // VmError raised by any row
};

}  // namespace ark
