"""Diagnostic: per-call wall times of the host-buffer path (load / table / graph / edge download)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sage2_b200 import api, synth

reads, k, cfg = bench.make_workload("cfg2")
bases, offsets = synth.concat(reads)
hb = torch.from_numpy(bases).pin_memory(); ho = torch.from_numpy(offsets).pin_memory()
gpu = api.Sage2Gpu(0)
buf = None
for it in range(8):
    t0 = time.perf_counter()
    gpu.load_reads_ptr(hb.data_ptr(), ho.data_ptr(), len(reads), k, device=False)
    t1 = time.perf_counter()
    gpu.build_hash_table()
    t2 = time.perf_counter()
    gpu.build_overlap_graph()
    t3 = time.perf_counter()
    if buf is None:
        buf = torch.empty(2 * gpu.counters()["n_edges"], dtype=torch.int64).pin_memory()
        t3 = time.perf_counter()
    gpu.edges_packed_into(buf.data_ptr(), buf.numel() // 2)
    t4 = time.perf_counter()
    tm = gpu.timers()
    print(it, "load %.1f table %.1f graph %.1f d2h %.1f ms | ingest %.1f sort %.1f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, tm["ingest"], tm["sort_reads"]), flush=True)
