"""ORACLE (test infrastructure, NOT product code): ctypes binding of the CPU restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from .binding import *  # noqa: F401,F403
