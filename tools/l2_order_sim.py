#!/usr/bin/env python
"""ANALYSIS AID for DESIGN.md section 6: LRU simulation of the search kernel's line requests (slot sectors + partner records)
for two read schedules -- id order (today) and min-hash order -- on a cfg2-shaped read set scaled down by `scale`, with the
126 MB L2 scaled by the same factor.   python tools/l2_order_sim.py [scale=20]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emul  # noqa: E402
from sage2_b200 import synth  # noqa: E402


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    so = "/tmp/l2_order_sim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so, os.path.join(ROOT, "tools", "l2_order_sim.cpp")])
    lib = C.CDLL(so)
    lib.l2sim.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    g = synth.random_genome(4_600_000 // scale, 4600)
    reads = synth.paired_reads(g, 150, 100, seed=4601, mu=450, sigma=30)
    b, off = synth.concat(reads)
    L = emul.lib()
    h = L.hemu_prepare(b.ctypes.data, off.ctypes.data, len(off) - 1, 63)
    sz = np.zeros(13, dtype=np.uint64)
    L.hemu_sizes(h, sz.ctypes.data)
    U, SW = int(sz[0]), int(sz[1])
    F, RC, ln = np.zeros(U * SW, np.uint64), np.zeros(U * SW, np.uint64), np.zeros(U, np.uint16)
    L.hemu_copy(h, F.ctypes.data, RC.ctypes.data, ln.ctypes.data, None, None, None, None, None, None)
    L.hemu_free(h)
    cache_lines = (126 << 20) // 128 // scale
    out = np.zeros(8, np.float64)
    inflight = max(1, 148 * 32 // scale)        # resident warps of the search kernel, scaled like the cache
    lib.l2sim(F.ctypes.data, RC.ctypes.data, ln.ctypes.data, U, SW, 63, cache_lines, inflight, out.ctypes.data)
    for mode, name in enumerate(("id order (today)", "min-hash order")):
        sr, sh, rr, rh = out[4 * mode:4 * mode + 4]
        print(f"{name:18s}: slot lines {sr:.3g} requests, {100 * sh / sr:.1f} % hits; partner records {rr:.3g} requests, {100 * rh / rr:.1f} % hits; "
              f"misses {sr - sh + rr - rh:.3g} lines")
    print(f"(U = {U}, cache = {cache_lines} lines = 126 MB / {scale}, {inflight} reads in flight)")


if __name__ == "__main__":
    main()
