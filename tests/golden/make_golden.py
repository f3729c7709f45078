#!/usr/bin/env python
"""Generate tests/golden/golden.json by running the UNMODIFIED reference (oracle/_ref/SAGE2).

Run in the build container only (needs /root/reference to have been compiled by
`make -C oracle ref`):   python tests/golden/make_golden.py

For every dataset in tests/datasets.py (plus cfg1 / cfg4mini from sage2_b200.synth) it writes the
interleaved FASTQ, runs `SAGE2 -f <fq> -k <k> -o <dir> -p g -s -M 3` with OMP_NUM_THREADS=1, and
records the md5 of the reference's own `.reads` and `.graph3` plus the counters from its log.
The `mixed` case also keeps the complete files as fixtures.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import re
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402
from oracle import oracle  # noqa: E402
from sage2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

COUNTERS = {
    "unique_reads": r"Number of unique reads: (\d+)",
    "good_reads": r"Good reads in file: (\d+)",
    "over_threshold": r"Number of hash elements over threshold: (\d+)",
    "contained_ext": r"Total contained by extension: (\d+)",
    "contained_size": r"Total contained by size: (\d+)",
    "left_to_explore": r"Total left to explore: (\d+)",
    "edges_inserted": r"Total edges inserted: (\d+)",
    "transitive_removed": r"Transitive edge removed: (\d+)",
}


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    names = [n for n in datasets.DATASETS if n not in datasets.NO_REFERENCE] + ["cfg1", "cfg4mini"]
    if len(sys.argv) > 1:
        names = sys.argv[1:]
    path = os.path.join(HERE, "golden.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    tmp = tempfile.mkdtemp(prefix="golden_")
    for name in names:
        reads, k = datasets.get(name) if name in datasets.DATASETS else synth.config(name)
        fq = os.path.join(tmp, name + ".fastq")
        synth.write_fastq(fq, reads)
        out = os.path.join(tmp, name)
        prefix = oracle.run_reference(fq, k, out, "g", max_step=3, threads=1, save=True, timeout=300)
        log = open(prefix + ".log").read().replace(",", "")
        entry = {"k": k, "n_reads": len(reads), "reads_md5": md5(prefix + ".reads"),
                 "graph3_md5": md5(prefix + ".graph3")}
        for key, pat in COUNTERS.items():
            m = re.search(pat, log)
            entry[key] = int(m.group(1)) if m else None
        gold[name] = entry
        if name in ("mixed",):
            for ext in (".reads", ".graph3"):
                with open(prefix + ext, "rb") as fi, gzip.GzipFile(os.path.join(HERE, name + ext + ".gz"), "wb", 9, mtime=0) as fo:
                    shutil.copyfileobj(fi, fo)
        print(name, entry, flush=True)
    json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
