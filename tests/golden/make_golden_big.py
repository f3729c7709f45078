#!/usr/bin/env python
"""Reference goldens at BENCHMARK size: cfg2, cfg3-40/60/90, cfg4 (BASELINE.json `configs` 2-4).

Run in the build container only:   python tests/golden/make_golden_big.py [names...] [--keep DIR]

For each workload it writes the interleaved FASTQ, runs the UNMODIFIED reference
(`oracle/_ref/SAGE2 -f <fq> -k <k> -o <dir> -p g -s -M 3`, main.cpp:37-132, all host cores; the
output is independent of the thread count) and records in
tests/golden/golden_big.json:
  * md5 of the reference's own `.reads` and `.graph3` (readLoader.cpp:270-287, overlapGraph.cpp:338-369),
  * the log counters,
  * `edges_digest` / `reads_digest`: order-sensitive 64-bit digests of the parsed files
    (tests/digest.py) which bench.py recomputes from the device arrays on every rank, so every
    benchmark number carries a bit-exact check without formatting 1 GB of text per step.
The files themselves stay in --keep DIR (default /tmp/golden_big) and are not committed.
"""
from __future__ import annotations

import json
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import digest  # noqa: E402
from make_golden import COUNTERS, md5  # noqa: E402
from oracle import oracle  # noqa: E402
from sage2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    args = sys.argv[1:]
    keep = "/tmp/golden_big"
    if "--keep" in args:
        i = args.index("--keep")
        keep = args[i + 1]
        del args[i:i + 2]
    names = args or ["cfg2", "cfg3-40", "cfg3-60", "cfg3-90", "cfg4"]
    path = os.path.join(HERE, "golden_big.json")
    os.makedirs(keep, exist_ok=True)
    threads = os.cpu_count()
    for name in names:
        gold = json.load(open(path)) if os.path.exists(path) else {}
        reads, k = synth.config(name)
        out = os.path.join(keep, name)
        prefix = os.path.join(out, "g")
        if not (os.path.exists(prefix + ".graph3") and os.path.exists(prefix + ".done")):
            fq = os.path.join(keep, name + ".fastq")
            synth.write_fastq(fq, reads)
            t0 = time.time()
            oracle.run_reference(fq, k, out, "g", max_step=3, threads=threads, save=True, timeout=6 * 3600)
            wall = time.time() - t0
            os.remove(fq)
            open(prefix + ".done", "w").write(f"{wall:.1f}\n")
        wall = float(open(prefix + ".done").read())
        log = open(prefix + ".log").read().replace(",", "")
        entry = {"k": k, "n_reads": len(reads), "reads_md5": md5(prefix + ".reads"),
                 "graph3_md5": md5(prefix + ".graph3"), "reference_wall_s": wall, "reference_threads": threads}
        for key, pat in COUNTERS.items():
            m = re.search(pat, log)
            entry[key] = int(m.group(1)) if m else None
        entry["edges_digest"], entry["n_edges"] = digest.graph3_file_digest(prefix + ".graph3")
        entry["reads_digest"] = digest.reads_file_digest(prefix + ".reads")
        gold[name] = entry
        print(name, entry, flush=True)
        json.dump(gold, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
