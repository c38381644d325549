// vm.h — row-expression bytecode shared by the planner (host) and the row kernels (device).
//
// A bound scalar expression (WHERE predicate, computed projection, aggregate argument) is compiled
// once per (query, input schema) into a short register program.  Every row evaluates the same
// program, so control flow inside the interpreter is warp-uniform.  The program travels in the
// kernel parameter block (`__grid_constant__`), never through global memory.
//
// Semantics restated from arrow-rs 55.2 / DataFusion 47 (third-party, not in /root/reference; call
// sites crates/arkflow-plugin/src/processor/sql.rs:126-129,197-203):
//   * Int64 + - * wrap (DataFusion BinaryExpr uses *_wrapping kernels unless fail_on_overflow);
//   * Int64 / and % by zero raise "Divide by zero error"; i64::MIN / -1 and i64::MIN % -1 raise arrow-arith's
//     ArithmeticOverflow (div_checked / mod_checked: checked_div / checked_rem return None for both);
//   * Float64 comparisons use IEEE-754 totalOrder (NaN above +Inf, -0.0 < +0.0, eq is bitwise);
//   * AND / OR are Kleene three-valued; a NULL predicate drops the row (FilterExec);
//   * CAST(Float64 AS BIGINT) truncates toward zero and errors on NaN / out-of-range.
#pragma once
#include <cstdint>

namespace ark {

enum VmOp : uint8_t {
  VM_NOP = 0,
  VM_LOAD_I64,   // dst <- column[a] (Int64), aux unused
  VM_LOAD_F64,   // dst <- column[a] (Float64)
  VM_LOAD_BOOL,  // dst <- column[a] (Boolean, bit-packed)
  VM_CONST,      // dst <- consts[aux] (raw 64-bit), valid
  VM_NULL,       // dst <- NULL
  VM_ADD_I64, VM_SUB_I64, VM_MUL_I64, VM_DIV_I64, VM_MOD_I64, VM_NEG_I64,
  VM_ADD_F64, VM_SUB_F64, VM_MUL_F64, VM_DIV_F64, VM_MOD_F64, VM_NEG_F64,
  VM_I64_TO_F64, VM_F64_TO_I64, VM_BOOL_TO_I64, VM_I64_TO_BOOL, VM_F64_TO_BOOL,
  VM_CMP_I64,    // dst <- a (cmp) b, aux = VmCmp
  VM_CMP_F64,    // totalOrder
  VM_CMP_BOOL,
  VM_CMP_STR_CONST,  // dst <- column[a] (Utf8/Binary) (cmp) string constant #b, aux = VmCmp
  VM_CMP_STR_COL,    // dst <- column[a] (cmp) column[b], aux = VmCmp
  VM_AND, VM_OR, VM_NOT,
  VM_IS_NULL, VM_IS_NOT_NULL,
  VM_IS_NULL_COL, VM_IS_NOT_NULL_COL,  // a = column index (works for var-len columns too)
};

enum VmCmp : int32_t { CMP_EQ = 0, CMP_NE, CMP_LT, CMP_LE, CMP_GT, CMP_GE };

enum VmError : int32_t { VMERR_NONE = 0, VMERR_DIV_ZERO = 1, VMERR_OVERFLOW = 2 /* i64::MIN / -1 */, VMERR_CAST = 3, VMERR_OVERFLOW_MOD = 4 /* i64::MIN % -1 */ };

struct VmInstr {
  uint8_t op;
  uint8_t dst;
  uint8_t a;
  uint8_t b;
  int32_t aux;
};

constexpr int VM_MAX_INSTR = 48;
constexpr int VM_MAX_CONST = 16;
constexpr int VM_MAX_REGS = 12;
constexpr int VM_MAX_STR_CONST = 4;
constexpr int VM_STR_CONST_BYTES = 64;

struct VmProgram {
  int32_t n_instr;    // 0 ⇒ "no expression" (predicate: every row passes)
  int32_t result_reg;
  VmInstr instr[VM_MAX_INSTR];
  uint64_t consts[VM_MAX_CONST];
  int32_t str_len[VM_MAX_STR_CONST];
  uint8_t str_bytes[VM_MAX_STR_CONST][VM_STR_CONST_BYTES];
};

// Device view of one input column (pointers already adjusted for the Arrow `offset`).
struct ColView {
  const void* data;         // values (fixed width), bit-packed bools, or string bytes base
  const int32_t* offsets;   // var-len only: offsets[row], offsets[row+1] index into data
  const uint8_t* validity;  // bitmap or nullptr
  int32_t validity_bit0;    // bit offset of row 0 inside validity
  int32_t data_bit0;        // Boolean only: bit offset of row 0 inside data
};

constexpr int MAX_COLS = 12;  // distinct input columns one kernel launch may reference

}  // namespace ark
