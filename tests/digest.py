"""Order-sensitive 64-bit digests of a unique-read table and of an edge list.

TEST / BENCH-CHECK INFRASTRUCTURE.  The same two digests are computed
  * here with numpy, from the reference's own text files (`.reads`, readLoader.cpp:270-287;
    `.graph3`, overlapGraph.cpp:338-369) or from arrays the oracle returns, and
  * on the device by `sage2gpu_digest` (sage2_b200/csrc/digest.cu) from the resident arrays,
so a benchmark run can prove on every rank that its result is the reference's, bit for bit,
without formatting and hashing a gigabyte of text per step.

Definitions (all arithmetic mod 2^64; mix = splitmix64 finaliser with the golden-ratio increment):
  edge i (0-based position in `.graph3` order: from ascending, then (to, type, overhang)):
      a = from << 32 | to ;  b = type << 48 | delta << 24 | delta_twin
      h_i = mix(mix(mix(i) + a) + b) ;           edges_digest = sum_i h_i + mix(E)
  unique read id (1-based rank, readLoader.cpp:225-235), forward strand packed 32 bases per
  64-bit word, first base in the top bits (codes A0 C1 G2 T3, utils.cpp:96-119), pad bits 0:
      h = mix(id) ; h = mix(h + (frequency << 16 | length)) ; for each of ceil(length/32) words: h = mix(h + word)
      reads_digest = sum_id h + mix(U)
"""
from __future__ import annotations

import numpy as np

_M = np.uint64
_G = _M(0x9E3779B97F4A7C15)
_C1 = _M(0xBF58476D1CE4E5B9)
_C2 = _M(0x94D049BB133111EB)


def mix(x):
    with np.errstate(over="ignore"):
        x = np.asarray(x, dtype=np.uint64) + _G
        x = (x ^ (x >> _M(30))) * _C1
        x = (x ^ (x >> _M(27))) * _C2
        return x ^ (x >> _M(31))


def edges_digest(frm, to, typ, delta, delta_twin, start: int = 0) -> int:
    """Digest contribution of edges at positions start.. (without the + mix(E) term)."""
    with np.errstate(over="ignore"):
        n = len(frm)
        i = np.arange(start, start + n, dtype=np.uint64)
        a = (np.asarray(frm, np.uint64) << _M(32)) | np.asarray(to, np.uint64)
        b = (np.asarray(typ, np.uint64) << _M(48)) | (np.asarray(delta, np.uint64) << _M(24)) | np.asarray(delta_twin, np.uint64)
        h = mix(mix(mix(i) + a) + b)
        return int(np.sum(h, dtype=np.uint64))


def finish(partial: int, count: int) -> int:
    return (partial + int(mix(np.uint64(count)))) & 0xFFFFFFFFFFFFFFFF


def edges_digest_total(edges) -> int:
    """`edges`: structured array with from/to/type/delta/delta_twin (oracle.EDGE_DT or api's)."""
    return finish(edges_digest(edges["from"], edges["to"], edges["type"], edges["delta"], edges["delta_twin"]), len(edges))


_CODE = np.full(256, 255, np.uint8)
for _c, _v in zip(b"ACGT", range(4)):
    _CODE[_c] = _v


def pack_words(ascii_rows: np.ndarray) -> np.ndarray:
    """(n, L) uint8 ASCII ACGT -> (n, ceil(L/32)) uint64, first base in the top bits."""
    n, L = ascii_rows.shape
    W = (L + 31) // 32
    codes = np.zeros((n, W * 32), np.uint64)
    codes[:, :L] = _CODE[ascii_rows]
    sh = (_M(62) - _M(2) * np.arange(32, dtype=np.uint64))
    return np.bitwise_or.reduce(codes.reshape(n, W, 32) << sh[None, None, :], axis=2)


def reads_digest(first_id: int, freq, length, words) -> int:
    """Reads of ONE length; `words` (n, ceil(length/32)).  Partial sum (no + mix(U))."""
    with np.errstate(over="ignore"):
        n = len(freq)
        ids = np.arange(first_id, first_id + n, dtype=np.uint64)
        h = mix(ids)
        h = mix(h + ((np.asarray(freq, np.uint64) << _M(16)) | np.asarray(length, np.uint64)))
        for w in range(words.shape[1]):
            h = mix(h + words[:, w])
        return int(np.sum(h, dtype=np.uint64))


def reads_digest_ragged(ids, freq, length, rows) -> int:
    """Partial digest for reads given as a list of ASCII byte strings (any lengths)."""
    ids = np.asarray(ids, np.uint64)
    freq = np.asarray(freq, np.uint64)
    length = np.asarray(length, np.int64)
    total = 0
    with np.errstate(over="ignore"):
        for L in np.unique(length):
            sel = np.nonzero(length == L)[0]
            mat = np.frombuffer(b"".join(rows[i] for i in sel), np.uint8).reshape(len(sel), int(L))
            words = pack_words(mat)
            h = mix(ids[sel])
            h = mix(h + ((freq[sel] << _M(16)) | _M(int(L))))
            for w in range(words.shape[1]):
                h = mix(h + words[:, w])
            total += int(np.sum(h, dtype=np.uint64))
    return total & 0xFFFFFFFFFFFFFFFF


def graph3_file_digest(path: str, chunk: int = 4_000_000):
    """(edges_digest, E) of a `.graph3` file: 3 header lines, then two records per undirected edge
    (`from to type 1 delta 0 0`, blank line, the twin, blank line)."""
    import pandas as pd
    total, pos = 0, 0
    it = pd.read_csv(path, sep="\t", header=None, skiprows=3, skip_blank_lines=True, chunksize=2 * chunk,
                     dtype=np.int64, engine="c")
    for df in it:
        a = df.to_numpy()
        assert len(a) % 2 == 0
        e, t = a[0::2], a[1::2]
        assert np.array_equal(e[:, 0], t[:, 1]) and np.array_equal(e[:, 1], t[:, 0])
        total += edges_digest(e[:, 0], e[:, 1], e[:, 2], e[:, 4], t[:, 4], start=pos)
        pos += len(e)
    return finish(total, pos), pos


def reads_file_digest(path: str, chunk: int = 1_000_000) -> int:
    """reads_digest of a `.reads` file: U, then `frequency length forward reverse` per read."""
    import pandas as pd
    total, first = 0, 1
    it = pd.read_csv(path, sep="\t", header=None, skiprows=1, usecols=[0, 1, 2], chunksize=chunk,
                     dtype={0: np.int64, 1: np.int64, 2: str}, engine="c")
    for df in it:
        freq, length = df[0].to_numpy(), df[1].to_numpy()
        rows = [s.encode() for s in df[2]]
        total += reads_digest_ragged(np.arange(first, first + len(rows)), freq, length, rows)
        first += len(rows)
    return finish(total, first - 1)
