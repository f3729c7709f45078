#!/usr/bin/env python
"""Step-6 mapping (sage2gpu_map_reads = ReadLoader::getIdOfRead for a batch) at cfg2: every input read
mapped back to its id.  Prints one JSON line: device-resident kernel time, the call with host buffers (H2D of the
reads + D2H of the ids inside), and the oracle's restatement of getIdOfRead on one host core for a sample."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sage2_b200 import api, synth  # noqa: E402


def main():
    import torch
    reads, k = synth.config("cfg2")
    b, off = synth.concat(reads)
    n = len(off) - 1
    g = api.Sage2Gpu(0)
    g.load_reads(b, off, k)
    hb, ho = torch.from_numpy(b).pin_memory(), torch.from_numpy(off).pin_memory()
    db, do = hb.cuda(), ho.cuda()
    ids = torch.empty(n, dtype=torch.int64).pin_memory()
    good = torch.empty(n, dtype=torch.uint8).pin_memory()
    ms_dev, ms_host = [], []
    for it in range(8):
        ms_dev.append(g.map_reads_ptr(db.data_ptr(), do.data_ptr(), n, ids.data_ptr(), good.data_ptr(), True))
    for it in range(5):
        t0 = time.perf_counter()
        g.map_reads_ptr(hb.data_ptr(), ho.data_ptr(), n, ids.data_ptr(), good.data_ptr(), False)
        ms_host.append((time.perf_counter() - t0) * 1e3)
    U = g.counters()["unique_reads"]
    assert bool((ids != 0).all())
    line = {"workload": "cfg2: every input read mapped to its id", "reads": n, "unique_reads": U,
            "kernel_ms": min(ms_dev[3:]), "reads_per_s_device_resident": n / (min(ms_dev[3:]) / 1e3),
            "host_buffers_ms": min(ms_host[1:]), "reads_per_s_host_buffers": n / (min(ms_host[1:]) / 1e3)}
    if "--cpu" in sys.argv:
        from oracle import oracle
        ns = 200_000
        o = oracle.OracleRun(b, off, k, threads=0)
        t0 = time.perf_counter()
        o.map_reads(b[:off[ns]], off[:ns + 1], k)
        dt = time.perf_counter() - t0
        line["cpu_port_reads_per_s_1_core"] = ns / dt
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
