"""Stage times on error-containing reads (real data has them; the BASELINE configs are error-free).
python tools/err_profile.py [genome_bp] [err_rate]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sage2_b200 import api, synth
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
err = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
g = synth.random_genome(G, 77)
reads = synth.paired_reads(g, 150, 100, seed=78, mu=450, sigma=30, err_rate=err)
b, off = synth.concat(reads)
gpu = api.Sage2Gpu(0)
for it in range(3):
    t = time.perf_counter()
    gpu.run_steps123(b, off, 63)
    dt = time.perf_counter() - t
c = gpu.counters()
print(f"G={G} err={err} reads={len(reads)} wall={dt*1e3:.1f} ms", {k: round(v, 2) for k, v in gpu.timers().items()})
print({k: c[k] for k in ("unique_reads", "slow_path_reads", "contained_ext", "left_to_explore", "candidates_c", "edges_inserted_c", "transitive_removed", "n_edges", "compare_calls")})
