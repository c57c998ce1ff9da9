"""CPU oracle for the MLS-MPM substep -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package.  The product (mpm_flip98a_b200/) never does.
"""
