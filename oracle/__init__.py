"""TEST INFRASTRUCTURE ONLY: CPU oracle for the complogic NAND-VM path (see clg_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product package (complogic_b200) never does.
"""
