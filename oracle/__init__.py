"""TEST INFRASTRUCTURE: CPU restatement (kuiper_oracle.c) and the compiled reference
(oracle/_ref).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product never does."""
