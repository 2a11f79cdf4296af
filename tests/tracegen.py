"""Shim: the trace generators live in workload/tracegen.py (shared by tests and bench.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workload.tracegen as _m  # noqa: E402

sys.modules[__name__] = _m
