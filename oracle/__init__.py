"""CPU oracle for the nnest hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this package -- and there only as the checker (or the timed CPU baseline), never as part
of the product path.  The product package `nnest_b200` must not import it.
"""
