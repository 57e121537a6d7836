"""CPU oracle for the FOTO / GN path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product path never does.  Parity status: pinned against
golden vectors generated from the unmodified reference (tests/golden/make_golden.py).
"""
from .oracle import *  # noqa: F401,F403
