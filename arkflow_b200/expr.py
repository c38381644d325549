"""Mirror of crates/arkflow-plugin/src/expr/mod.rs and of the Temporary / Resource types the `sql`
processor's `temporary_list` enrichment uses (crates/arkflow-core/src/temporary/mod.rs:36-41,
crates/arkflow-plugin/src/processor/sql.rs:151-186).  Expression evaluation runs on the GPU through
`ark_expr_evaluate`; the Temporary store itself (redis in the reference) is the caller's object."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import pyarrow as pa

from . import _lib as L
from . import arrow_ffi as F


class ColumnarValue:
    """datafusion::logical_expr::ColumnarValue: Array(values) | Scalar(value)."""

    def __init__(self, kind: str, array: pa.Array):
        self.kind, self.array = kind, array  # Scalar keeps a one-element array

    @staticmethod
    def scalar_utf8(value: str) -> "ColumnarValue":
        return ColumnarValue("Scalar", pa.array([value], pa.utf8()))

    def is_scalar(self) -> bool:
        return self.kind == "Scalar"

    def scalar_value(self):
        assert self.kind == "Scalar"
        return self.array[0].as_py()


def evaluate_expr(expr_str: str, batch: pa.RecordBatch) -> ColumnarValue:
    """expr::evaluate_expr (expr/mod.rs:92-122).  Raises ArkError with the planner's / evaluator's message."""
    from .processor import _check

    arr, sch = F.export_record_batch(batch)
    out_arr, out_sch = L.ArrowArray(), L.ArrowSchema()
    is_scalar = C.c_int(0)
    try:
        status = L.lib().ark_expr_evaluate(expr_str.encode(), C.byref(arr), C.byref(sch), C.byref(out_arr), C.byref(out_sch), C.byref(is_scalar))
    finally:
        F.release_schema(sch)
        F.release_array(arr)
    _check(status)
    rb = F.import_record_batch(out_arr, out_sch)
    return ColumnarValue("Scalar" if is_scalar.value else "Array", rb.column(0))


class EvaluateResult:
    """expr/mod.rs:37-49: Scalar(T) | Vec(Vec<T>); get(i) returns the scalar for every i."""

    def __init__(self, kind: str, value):
        self.kind, self.value = kind, value

    def get(self, i: int):
        if self.kind == "Scalar":
            return self.value
        return self.value[i] if 0 <= i < len(self.value) else None


class Expr:
    """expr/mod.rs:30-35 — serde-tagged {"type": "expr", "expr": "..."} | {"type": "value", "value": ...}."""

    def __init__(self, kind: str, payload):
        self.kind, self.payload = kind, payload

    @staticmethod
    def from_config(cfg: dict) -> "Expr":
        from .processor import ArkError

        t = cfg.get("type") if isinstance(cfg, dict) else None
        if t == "expr" and isinstance(cfg.get("expr"), str):
            return Expr("Expr", cfg["expr"])
        if t == "value" and "value" in cfg:
            return Expr("Value", cfg["value"])
        raise ArkError(L.ARK_ERR_SERIALIZATION, "invalid Expr: expected {type: expr, expr: …} or {type: value, value: …}")

    def evaluate_expr(self, batch: pa.RecordBatch) -> EvaluateResult:
        """Expr<String>::evaluate_expr (expr/mod.rs:51-90)."""
        from .processor import ArkError

        if self.kind == "Value":
            return EvaluateResult("Scalar", self.payload)
        try:
            cv = evaluate_expr(self.payload, batch)
        except ArkError as e:
            raise ArkError(L.ARK_ERR_PROCESS, f"Failed to evaluate expression: {e.message}")
        if cv.kind == "Array":
            if cv.array.type != pa.utf8():
                raise ArkError(L.ARK_ERR_PROCESS, "Failed to evaluate expression")
            return EvaluateResult("Vec", [s for s in cv.array.to_pylist() if s is not None])  # filter_map drops NULLs
        if cv.array.type != pa.utf8():
            raise ArkError(L.ARK_ERR_PROCESS, f"Unsupported scalar type: {_df_type_name(cv.array.type)}")
        v = cv.scalar_value()
        if v is None:
            raise ArkError(L.ARK_ERR_PROCESS, "Null string value")
        return EvaluateResult("Scalar", v)


def _df_type_name(t: pa.DataType) -> str:
    return {pa.int64(): "Int64", pa.float64(): "Float64", pa.bool_(): "Boolean", pa.utf8(): "Utf8", pa.binary(): "Binary"}.get(t, str(t))


class Temporary:
    """trait Temporary (core/temporary/mod.rs:36-41)."""

    def connect(self) -> None:
        pass

    def get(self, keys: list[ColumnarValue]):  # -> Optional[MessageBatch]
        raise NotImplementedError

    def close(self) -> None:
        pass


class Resource:
    """core/lib.rs Resource: the temporaries by name and the input names seen while building the stream."""

    def __init__(self, temporary: Optional[dict] = None, input_names: Optional[list] = None):
        self.temporary = dict(temporary or {})
        self.input_names = list(input_names or [])
