"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/vct_oracle.h).  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
