"""CPU oracles for the Segmenter forward path.  TEST INFRASTRUCTURE ONLY.

Nothing under sylber_b200/ may import this package; only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py do.
"""
