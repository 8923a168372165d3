"""Synthetic workloads of the shape BASELINE.json names (frames, ult_logger logs, weights).

Neither product nor oracle: `bench.py`, `tests/` and `__graft_entry__.smoke()` import this package
to build seeded inputs; `playaid_core_b200` never does.
"""
