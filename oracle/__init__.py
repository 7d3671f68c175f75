"""oracle/ — CPU restatements of the reference algorithms.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(text2human_b200/) never imports it.
"""
