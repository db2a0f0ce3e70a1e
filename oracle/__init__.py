"""CPU oracle for the cloud_merger merge hot path -- TEST INFRASTRUCTURE ONLY (parity unpinned, see cm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
