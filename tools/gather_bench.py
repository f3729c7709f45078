"""Random-sector gather ceiling of this GPU (sage2gpu_measure_gather): GB/s and gathers/s by granule, load shape and footprint."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sage2_b200 import api
g = api.Sage2Gpu(0)
for gran, mode in ((16, 0), (32, 0), (32, 1), (64, 0), (64, 1), (64, 2)):
    for fp_mb in (48, 320, 8192):
        gbs = g.measure_gather(fp_mb << 20, gran, 1 << 28, mode)
        print(f"gather granule={gran}B mode={mode} footprint={fp_mb}MiB {gbs:8.1f} GB/s {gbs / gran:6.2f} G gathers/s", flush=True)
