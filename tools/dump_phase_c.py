import os, sys
sys.path.insert(0, "/root/repo")
os.environ["SAGE2GPU_DUMP_PHASE_C"] = "/root/repo/gpurun_out/phasec_cfg4.bin"
from sage2_b200 import api, synth
reads, k = synth.config("cfg4")
b, off = synth.concat(reads)
g = api.Sage2Gpu(0)
g.run_steps123(b, off, k)
print(g.counters(), g.timers())
