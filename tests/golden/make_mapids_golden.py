#!/usr/bin/env python
"""Generate tests/golden/mapids.json: ReadLoader::getIdOfRead of the UNMODIFIED reference (oracle/_ref/ref_mapids,
our main() around the reference's own classes, `make -C oracle ref`) for the query reads of
tests/datasets.map_queries().  Run in the build container only:   python tests/golden/make_mapids_golden.py"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402
from sage2_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["clean", "mixed", "varlen_err", "tandem", "k70"]


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_mapids")
    gold = {}
    tmp = tempfile.mkdtemp(prefix="mapids_")
    for name in NAMES:
        reads, k = datasets.get(name)
        queries, _ = datasets.map_queries(name)
        fq, qq = os.path.join(tmp, name + ".fastq"), os.path.join(tmp, name + ".q.fastq")
        synth.write_fastq(fq, reads)
        synth.write_fastq(qq, queries)
        out = subprocess.check_output([exe, fq, qq, str(k)], env=dict(os.environ, OMP_NUM_THREADS="1")).decode().split()
        ids = [0 if t == "bad" else int(t) for t in out]
        bad = [1 if t == "bad" else 0 for t in out]
        gold[name] = {"k": k, "n_queries": len(queries), "n_answers": len(out),
                      "md5": hashlib.md5("\n".join(out).encode()).hexdigest(),
                      "positive": sum(1 for i in ids if i > 0), "negative": sum(1 for i in ids if i < 0),
                      "absent": sum(1 for i, b in zip(ids, bad) if i == 0 and not b), "bad": sum(bad), "first": out[:64]}
        print(name, {kk: vv for kk, vv in gold[name].items() if kk != "first"}, flush=True)
    json.dump(gold, open(os.path.join(HERE, "mapids.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
