"""CPU oracle for the two-electron hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (mcmurchie-davidson_b200/) never does.
"""
