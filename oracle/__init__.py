"""CPU oracle of the disconnected-loop hot path — TEST INFRASTRUCTURE ONLY (see mugiq_oracle.cpp).
Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
