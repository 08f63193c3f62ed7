"""ORACLE — test infrastructure only.

CPU restatements of the reference's algorithms for the hot path. Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this package; the product path
(`argus_b200/`) never does and fails loudly when libargus_b200.so is missing.
"""
