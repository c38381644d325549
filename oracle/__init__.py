"""oracle/ — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

May be imported by: tests/, __graft_entry__.smoke(), bench.py's `cpu_baseline` leg and
`bench.py --impl reference`.  The product (arkflow_b200/) never imports it and has no CPU fallback.

Parity status: the reference (Rust on un-vendored DataFusion 47 / arrow-rs 55.2) cannot be built in
this image (no cargo/rustc, no network), so this is a *restatement*; it is pinned against every
numeric/behavioural assertion the reference's own tests hold for this path
(tests/golden/reference_pins.json, built by tests/golden/make_golden.py, checked by
tests/test_oracle_golden.py).  SUM/AVG values, GROUP BY contents and JOIN contents are
"parity unpinned" by the reference and say so in DESIGN.md.
"""
