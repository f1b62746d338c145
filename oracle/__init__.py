"""CPU oracle (test infrastructure). See oracle/recom_oracle.c for what it restates.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this package. The product (recom_b200) never does.
"""
