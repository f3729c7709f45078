"""One gather microbenchmark configuration (for ncu): python tools/gather_one.py <granule> <mode> <footprint_MiB>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sage2_b200 import api
g = api.Sage2Gpu(0)
gran, mode, fp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
print(gran, mode, fp, g.measure_gather(fp << 20, gran, 1 << 26, mode))
