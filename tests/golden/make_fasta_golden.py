#!/usr/bin/env python
"""Golden contig / scaffold FASTA of STRAIGHT-THROUGH runs of the unmodified reference (all seven steps, main.cpp:11-297).

Run in the build container only:   python tests/golden/make_fasta_golden.py [names...]

`SAGE2 -f <fq> -k <k> -o <dir> -p g` with OMP_NUM_THREADS=1 (step 6 mutates shared lists inside an `omp for`,
matePair.cpp:168-237, so only a 1-thread run is reproducible by construction).  tests/test_gpu_shim.py runs
oracle/_ref/SAGE2_gpu -- the same main.cpp with steps 1-3 on the GPU through sage2_b200/host/sage2gpuShim.cpp -- on the
same input and compares the md5 of every output file recorded here (tests/golden/fasta_golden.json).
"""
from __future__ import annotations

import hashlib
import json
import os
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
FILES = ("g_contig.fasta", "g_scaffold.fasta", "g_contig.gdl", "g_scaffold.gdl")


def main():
    names = sys.argv[1:] or ["cfg1", "cfg4mini", "rep", "k70", "tandem"]
    path = os.path.join(HERE, "fasta_golden.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    tmp = tempfile.mkdtemp(prefix="fasta_golden_")
    cwd = os.getcwd()
    os.chdir(tmp)                     # step 5 writes input.cs2 / output.cs2 into the working directory
    for name in names:
        reads, k = datasets.get(name) if name in datasets.DATASETS else synth.config(name)
        fq = os.path.join(tmp, name + ".fastq")
        synth.write_fastq(fq, reads)
        out = os.path.join(tmp, name)
        try:
            oracle.run_reference(fq, k, out, "g", max_step=7, threads=1, save=False, timeout=3600)
        except Exception as ex:       # noqa: BLE001 - some tiny sets do not survive the reference's own steps 4-7
            print(name, "reference failed:", ex, flush=True)
            continue
        entry = {"k": k, "n_reads": len(reads)}
        for f in FILES:
            entry[f] = hashlib.md5(open(os.path.join(out, f), "rb").read()).hexdigest()
            entry[f + ".bytes"] = os.path.getsize(os.path.join(out, f))
        gold[name] = entry
        print(name, entry, flush=True)
    os.chdir(cwd)
    json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
