"""CPU oracle for the ngsDist hot path: TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; nothing under ngsdist_b200/ does.  See oracle/ngsdist_oracle.c for the restatement and its parity
status, oracle/Makefile for how oracle/_ref/ngsDist (the unmodified reference) is built.
"""
from .oracle import *  # noqa: F401,F403
