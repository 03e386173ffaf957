"""CPU parity oracle for proxtv_b200 -- TEST INFRASTRUCTURE ONLY.

Never import this from proxtv_b200/ (the product).  Allowed importers: tests/, __graft_entry__.smoke(),
bench.py (cpu_baseline / --impl reference legs).  See oracle/tv_oracle.c for the parity-pin statement.
"""
from .oracle import *  # noqa: F401,F403
