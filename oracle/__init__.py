"""CPU oracle for the Lux hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (lux_b200) never does.  See oracle/lux_oracle.c for the
reference file:line each function restates and for the parity-pin status.
"""
from .lux_oracle import *  # noqa: F401,F403
