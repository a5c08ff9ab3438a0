"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot path (eps696/aphantasia) used as the checker for the CUDA
path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (aphantasia_b200/) never does.
"""
