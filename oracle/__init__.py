"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot-path arithmetic (C for the CUDA kernels, numpy/torch
for the python-level pieces) plus a ctypes binding of the reference's own kernels compiled
unmodified for sm_100a (oracle/_ref).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this package; nothing under nesie_b200/ does.
"""
