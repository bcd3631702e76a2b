"""CPU oracle for the JMODT hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke()
may import this package.  The product (jmodt_b200/) never does and fails loudly if its
CUDA library is missing.
"""
