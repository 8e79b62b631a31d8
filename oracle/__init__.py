"""TEST INFRASTRUCTURE ONLY: CPU oracle for the photic per-pixel inversion (see photic_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (photic_b200/) never does.
"""
