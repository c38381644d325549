"""Randomised differential test of the expression VM: seeded random scalar expressions over Int64 /
Float64 / Boolean / Utf8 columns with NULLs and edge values, evaluated as a projection and as a WHERE
clause by the CUDA path and by oracle/sql_oracle.py.  Both sides must agree on the values (bit-exact,
NaN-aware) or on the error class (division by zero, cast overflow, planning errors)."""
import math
import random

import numpy as np
import pyarrow as pa
import pytest

from arkflow_b200.processor import ArkError, MessageBatch, SqlProcessor
from oracle.sql_oracle import OracleError, sql_process

pytestmark = pytest.mark.gpu

N_ROWS = 4099


def make_batch(seed):
    rng = np.random.default_rng(seed)
    n = N_ROWS

    def with_nulls(vals, typ, p=0.1):
        mask = rng.random(n) < p
        return pa.array([None if m else v for v, m in zip(vals, mask)], typ)

    small = rng.integers(-20, 21, n).tolist()
    edge = [0, 1, -1, 2**31, -(2**31), 2**62, -(2**62), 2**63 - 1, -(2**63), 7, -7, 3]
    i2 = [edge[k % len(edge)] if k % 17 == 0 else int(v) for k, v in enumerate(rng.integers(-1000, 1000, n))]
    fvals = (rng.random(n) * 200 - 100).tolist()
    fedge = [0.0, -0.0, float("nan"), float("inf"), -float("inf"), 1e300, -1e300, 5e-324, 0.5, -2.5, 9.3e18, -9.3e18]
    f2 = [fedge[k % len(fedge)] if k % 13 == 0 else v for k, v in enumerate((rng.random(n) * 10).tolist())]
    words = ["", "a", "temp_1", "temp_2", "zebra", "Zebra", "ünï", "temp_10", "b"]
    return pa.record_batch({
        "i1": with_nulls(small, pa.int64()),
        "i2": with_nulls(i2, pa.int64(), 0.05),
        "f1": with_nulls(fvals, pa.float64()),
        "f2": with_nulls(f2, pa.float64(), 0.05),
        "b1": with_nulls((rng.random(n) < 0.5).tolist(), pa.bool_()),
        "s1": with_nulls([words[k] for k in rng.integers(0, len(words), n)], pa.utf8()),
    })


class Gen:
    def __init__(self, seed):
        self.r = random.Random(seed)

    def num(self, depth):
        r = self.r
        if depth <= 0 or r.random() < 0.3:
            return r.choice(["i1", "i2", "f1", "f2", str(r.randint(-5, 5)), str(r.randint(0, 3)), "2.5", "0.0", "10", "-1"])
        k = r.random()
        if k < 0.55:
            return f"({self.num(depth - 1)} {r.choice(['+', '-', '*', '/', '%'])} {self.num(depth - 1)})"
        if k < 0.65:
            return f"(- {self.num(depth - 1)})"
        if k < 0.8:
            return f"CAST({self.num(depth - 1)} AS DOUBLE)"
        if k < 0.9:
            return f"CAST({self.num(depth - 1)} AS BIGINT)"
        return f"CAST({self.boolean(depth - 1)} AS BIGINT)"

    def boolean(self, depth):
        r = self.r
        if depth <= 0 or r.random() < 0.2:
            return r.choice(["b1", "i1 >= 0", "f1 < 10", "s1 = 'temp_1'", "i1 IS NULL", "f2 IS NOT NULL", "true", "false"])
        k = r.random()
        if k < 0.4:
            return f"({self.num(depth - 1)} {r.choice(['=', '!=', '<', '<=', '>', '>='])} {self.num(depth - 1)})"
        if k < 0.55:
            return f"(s1 {r.choice(['=', '!=', '<', '<=', '>', '>='])} '{r.choice(['temp_1', 'a', 'temp_10', '', 'zz'])}')"
        if k < 0.8:
            return f"({self.boolean(depth - 1)} {r.choice(['AND', 'OR'])} {self.boolean(depth - 1)})"
        if k < 0.9:
            return f"(NOT {self.boolean(depth - 1)})"
        return f"({self.num(depth - 1)} IS {r.choice(['', 'NOT '])}NULL)"


def outcome_oracle(rb, q):
    try:
        return "ok", sql_process(rb, q)
    except OracleError as e:
        return e.kind, None


def outcome_gpu(rb, q):
    try:
        r = SqlProcessor({"query": q}).process(MessageBatch.new_arrow(rb))
        return "ok", (None if r.is_none() else r.batches[0].record_batch)
    except ArkError as e:
        return e.kind, None


def same_column(a: pa.Array, b: pa.Array) -> bool:
    if a.type != b.type or len(a) != len(b):
        return False
    if a.type == pa.float64():
        va, vb = np.asarray(a.is_valid()), np.asarray(b.is_valid())
        if not np.array_equal(va, vb):
            return False
        x = a.fill_null(0.0).to_numpy(zero_copy_only=False).view(np.uint64)[va]
        y = b.fill_null(0.0).to_numpy(zero_copy_only=False).view(np.uint64)[vb]
        xn = np.isnan(x.view(np.float64))
        return bool(np.array_equal(xn, np.isnan(y.view(np.float64))) and np.array_equal(x[~xn], y[~xn]))  # any NaN payload = NaN
    return a.equals(b)


@pytest.mark.parametrize("seed", range(12))
def test_random_expressions(gpu, seed):
    rb = make_batch(1000 + seed)
    g = Gen(seed)
    agree_ok = 0
    for k in range(40):
        depth = 1 + k % 4
        queries = [f"SELECT {g.num(depth)} AS r, i1 FROM flow", f"SELECT {g.boolean(depth)} AS r FROM flow",
                   f"SELECT i2, s1 FROM flow WHERE {g.boolean(depth)}"]
        for q in queries:
            ko, wo = outcome_oracle(rb, q)
            kg, wg = outcome_gpu(rb, q)
            assert ko == kg, (q, ko, kg)
            if ko != "ok":
                continue
            agree_ok += 1
            assert (wo is None) == (wg is None), q
            if wo is None:
                continue
            assert wg.schema.names == wo.schema.names, q
            assert wg.num_rows == wo.num_rows, (q, wg.num_rows, wo.num_rows)
            for name in wo.schema.names:
                assert same_column(wg.column(name), wo.column(name)), (q, name)
    assert agree_ok >= 40  # most expressions evaluate; the rest agree on the error class
