"""CPU checkers for smvs_b200. TEST INFRASTRUCTURE ONLY: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product path (smvs_b200/) must not."""
