"""CPU oracle for the MoE hot path — TEST INFRASTRUCTURE ONLY (parity unpinned; see
oracle/moe_oracle.py).  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports it."""
