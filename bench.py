#!/usr/bin/env python
"""bench.py -- reads/s through SAGE2's overlap-graph build (reference steps 1-3) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one complete pass of the hot path over one batch of synthetic reads: ingest (filter,
canonicalise, 2-bit pack), sort + dedupe, prefix/suffix table, phase A search + extension state
machine, phase B, phase-C candidates + host walk, canonical edge sort.  The workload is BASELINE.json
config #2 (4.6 Mbp random genome, 150 bp paired-end, 100x, -k 63): 3,066,666 input reads per step.

  value   : whole-job input reads/s with the ASCII reads already resident in HBM.
  e2e     : the same through the C ABI with HOST (pinned) buffers: H2D of the reads and D2H of the
            edge list are inside the timed region.
  roofline: the phase-A search kernel against the measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference: the UNMODIFIED reference's steps 1-3 (oracle/_ref/ref_steps123,
            built from /root/reference in the build container) on the box's host cores, on a bounded
            cfg2-shaped sample.

N > 1 (torchrun, one rank per GPU): ONE cfg2 read set for the whole job, total work fixed, so scaling is
"strong".  Two layouts, the headline is the faster one at this size:
  --table replicated (default): reads and table on every GPU, phase A partitioned by read id, one NCCL exchange;
  --table sharded: every GPU holds one key-hash shard of the table, the window probes are routed to their owners
            through peer-memory mailboxes (--exchange p2p: kernel stores over NVLink) or NCCL all-to-all.
The other layout is timed on the same reads and reported as `alt_table`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/sec through exact-overlap graph build (SAGE2 steps 1-3)"
UNIT = "reads/s"


def make_workload(name: str, seed_shift: int = 0, genome_size: int | None = None):
    from sage2_b200 import synth
    if name != "cfg2":
        reads, k = synth.config(name)
        return reads, k, {"workload": name}
    G = genome_size or 4_600_000
    g = synth.random_genome(G, 4600 + 7919 * seed_shift)
    reads = synth.paired_reads(g, 150, 100, seed=4601 + 7919 * seed_shift, mu=450, sigma=30)
    return reads, 63, {"workload": "cfg2: synthetic 4.6 Mbp random genome, 150 bp paired-end, 100x, -k 63",
                       "genome_bp": G, "read_len": 150, "coverage": 100, "k": 63}


# ---- clocks ------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons during the timed region, read through NVML in a background thread
    (a polling `nvidia-smi -lms` subprocess was measured to stall kernel launches for tens of ms)."""

    def __init__(self, index: int, period_s: float = 0.1):
        self.rows, self.stop_flag, self.t, self.err = [], threading.Event(), None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.period = period_s
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:      # noqa: BLE001 - any NVML problem just means "no clock record"
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            self.stop_flag.wait(self.period)

    def stop(self) -> dict:
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + str(self.err)]}
        self.stop_flag.set()
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        sm = [c for c, _ in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max, "reasons": reasons,
                "samples": len(sm), "source": "NVML"}


# ---- the reference on the host cores -------------------------------------------------------------

def reference_run(reads, k, tmpdir: str, threads: int | None = None) -> dict:
    """One run of the unmodified reference's steps 1-3 (oracle/_ref/ref_steps123) on `reads`."""
    from oracle import oracle
    from sage2_b200 import synth
    fq = os.path.join(tmpdir, "sample.fastq")
    synth.write_fastq(fq, reads)
    env = dict(os.environ)
    # all host cores unless told otherwise (torchrun exports OMP_NUM_THREADS=1 to its ranks)
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
    if os.access(oracle.REF_STEPS, os.X_OK):
        out = subprocess.run([oracle.REF_STEPS, fq, str(k)], env=env, check=True, capture_output=True, text=True).stdout
        d = json.loads(out.strip().splitlines()[-1])
        d["kind"] = "reference"
        return d
    # the reference did not travel: time the C port instead
    b, off = synth.concat(reads)
    t = time.perf_counter()
    o = oracle.OracleRun(b, off, k, threads or 0)
    dt = time.perf_counter() - t
    return {"input_reads": len(reads), "unique_reads": o.U, "edges": o.n_edges, "threads": threads or os.cpu_count(),
            "t_steps123": dt, "kind": "port"}


def sample_workload(target_seconds: float, tmpdir: str):
    """A cfg2-shaped sample (same read length, coverage, k; smaller genome) sized for ~target_seconds."""
    probe_reads, k, _ = make_workload("cfg2", genome_size=150_000)
    d = reference_run(probe_reads, k, tmpdir)
    rate = d["input_reads"] / max(d["t_steps123"], 1e-3)
    want_reads = min(3_066_666, max(100_000, int(rate * target_seconds)))
    G = max(150_000, min(4_600_000, int(want_reads * 150 / 100)))
    return make_workload("cfg2", genome_size=G)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tmp = tempfile.mkdtemp(prefix="sage2_ref_")
    per_step = max(4.0, 170.0 / max(1, args.steps + args.warmup))
    reads, k, cfg = sample_workload(min(25.0, per_step), tmp)
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = reference_run(reads, k, tmp)
        if i >= args.warmup:
            times.append(last["t_steps123"])
    ms = 1000.0 * float(np.mean(times))
    value = len(reads) / (ms / 1000.0)
    sample = f"cfg2-shaped sample: {cfg['genome_bp']} bp genome, {len(reads)} reads, steps 1-3 without FASTQ parse / text output"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "impl": "reference",
            "config": dict(cfg, workload=cfg["workload"], sample_reads=len(reads)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last.get("threads"), "kind": last["kind"], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_breakdown_s": {k2: v for k2, v in last.items() if k2.startswith("t_")}}
    print(json.dumps(line), flush=True)


# ---- our arm -----------------------------------------------------------------------------------

def algorithmic_bytes_phase_a(c: dict, read_len: int, n_slice: int | None = None) -> float:
    """SURVEY.md 8(d) restricted to the phase-A launch: stream the packed reads once, one 32-B sector
    per window probe, one sector-rounded packed partner read per gated comparison, 16 B of extension
    records out per read."""
    U, V, probes = (n_slice if n_slice is not None else c["unique_reads"]), c["compare_calls"], c["window_probes"]
    packed = (read_len + 3) // 4
    B = 32 * ((packed + 31) // 32)
    return float(U * packed + 32 * probes + B * V + 16 * U)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sage2_b200 import api, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # N > 1: ONE read set for the whole job (same seed on every rank); phase A is partitioned by read id and followed
    # by one NCCL exchange (sage2_b200/multi.py); reads and table are replicated
    from sage2_b200 import multi
    reads, k, cfg = make_workload(args.workload, seed_shift=0)
    n_reads = len(reads)
    read_len = int(reads.shape[1]) if isinstance(reads, np.ndarray) and reads.ndim == 2 else 0
    bases, offsets = synth.concat(reads)
    h_bases = torch.from_numpy(bases).pin_memory()
    h_off = torch.from_numpy(offsets).pin_memory()
    d_bases = h_bases.cuda(non_blocking=False)
    d_off = h_off.cuda(non_blocking=False)
    gpu = api.Sage2Gpu(local)
    stream = torch.cuda.ExternalStream(gpu.stream_ptr(), device=torch.device("cuda", local))

    dev = torch.device("cuda", local)
    comm = {"sent": 0, "h2d": 0}

    # the mailbox transport needs peer access between all GPUs of the job (NVLink / NVSwitch box); otherwise NCCL
    peers_ok = True
    if world > 1:
        try:
            peers_ok = torch.cuda.device_count() >= world and all(torch.cuda.can_device_access_peer(local, d) for d in range(world) if d != local)
        except Exception:
            peers_ok = False
    if world > 1:
        t_ok = torch.tensor([1 if peers_ok else 0], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        peers_ok = bool(t_ok.item())
    layout = {"sharded": args.table == "sharded"}
    xstats = {}

    def build_graph():
        if layout["sharded"]:       # key-hash shard per GPU, window probes routed to their owners by NCCL all-to-all
            gpu.build_hash_table_shard(rank, world)
            p2p = args.exchange == "p2p" and peers_ok      # peer-memory mailboxes (NVLink P2P stores) or NCCL all-to-all
            return multi.build_overlap_graph_sharded(gpu, rank, world, dev, batch_reads=min(args.batch_reads, 1 << 19) if p2p else args.batch_reads,
                                                     stats=xstats, p2p=p2p)
        gpu.build_hash_table()
        return multi.build_overlap_graph(gpu, rank, world, dev)

    def step_device():
        gpu.load_reads_ptr(d_bases.data_ptr(), d_off.data_ptr(), n_reads, k, device=True)
        comm["sent"] = build_graph()

    h_edges = {"buf": None}

    def step_host():
        # the call a user of the C ABI makes: host buffers in, edge list back in host memory
        if world == 1:
            gpu.load_reads_ptr(h_bases.data_ptr(), h_off.data_ptr(), n_reads, k, device=False)
            comm["h2d"] = int(bases.nbytes + offsets.nbytes)
        else:       # each rank moves 1/N of the input over PCIe, NVLink all-gather completes it
            # (on torch's own stream: pinned tensors must not be tied to the library's stream, which dies first)
            tb, to, comm["h2d"] = multi.upload_partitioned(h_bases, h_off, rank, world, dev)
            torch.cuda.current_stream(dev).synchronize()
            gpu.load_reads_ptr(tb.data_ptr(), to.data_ptr(), n_reads, k, device=True)
        build_graph()
        if h_edges["buf"] is None:
            h_edges["buf"] = torch.empty(2 * max(1, gpu.counters()["n_edges"]), dtype=torch.int64).pin_memory()
        gpu.edges_packed_into(h_edges["buf"].data_ptr(), h_edges["buf"].numel() // 2)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_step = []

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = api.kernel_launches()
        t0 = time.perf_counter()
        ev0.record(stream)
        stage = {}
        marks = [ev0]
        for _ in range(steps):
            fn()
            for kk, vv in gpu.timers().items():
                stage[kk] = stage.get(kk, 0.0) + vv
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record(stream)
        ev1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        dev_ms = ev0.elapsed_time(ev1)
        per_step.append([round(a.elapsed_time(b), 3) for a, b in zip(marks[:-1], marks[1:])])
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), api.kernel_launches() - l0, {a: b / steps for a, b in stage.items()}

    for _ in range(max(3, args.warmup)):
        step_device()
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("SAGE2_BENCH_NO_SAMPLER") else None
    dev_ms, wall_ms, launches, stage = timed(step_device, args.steps)
    counters = gpu.counters()
    clocks = sampler.stop() if sampler else None
    step_host()
    e2e_dev_ms, e2e_wall_ms, _, _ = timed(step_host, args.steps)
    n_edges = gpu.counters()["n_edges"]
    sharded = layout["sharded"]
    main_sent = comm["sent"]

    # N > 1: the other table layout on the same reads, device-resident, reported next to the headline
    alt = None
    if world > 1 and not args.no_alt_table:
        layout["sharded"] = not sharded
        try:
            xstats.clear()
            for _ in range(3):
                step_device()
            xstats.clear()
            a_ms, _, a_launches, a_stage = timed(step_device, args.steps)
            alt = {"table": "sharded" if layout["sharded"] else "replicated",
                   "exchange": (args.exchange if peers_ok else "nccl") if layout["sharded"] else "nccl",
                   "ms_per_step": a_ms / args.steps,
                   "value": n_reads / (a_ms / args.steps / 1000.0), "unit": UNIT, "sent_bytes_per_rank_and_step": comm["sent"],
                   "gpu_launches": a_launches, "stage_ms": a_stage,
                   "exchange_wall_ms_per_step": {kk: (vv / args.steps) for kk, vv in xstats.items()}}
        except Exception as ex:      # the headline above stands; the failure is reported, not hidden
            alt = {"table": "sharded" if layout["sharded"] else "replicated", "error": repr(ex)[:300]}
        layout["sharded"] = sharded
        comm["sent"] = main_sent

    def teardown():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        gpu.close()

    if rank != 0:
        teardown()
        return

    ms_per_step = dev_ms / args.steps
    value = n_reads / (ms_per_step / 1000.0)             # one job: every rank worked on the same n_reads
    e2e_value = n_reads / (e2e_dev_ms / args.steps / 1000.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_slice = min(counters["unique_reads"], -(-counters["unique_reads"] // world))      # rank 0's share of the reads
    abytes = algorithmic_bytes_phase_a(counters, read_len or counters["avg_len"], n_slice)
    traffic = None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "phase_a_traffic.json")))
        if world == 1 and prof.get("workload") == args.workload:
            traffic = float(prof["dram_bytes_per_launch"])
    except (OSError, ValueError, KeyError):
        pass
    gather = None
    if not args.no_gather:
        # the random-access ceiling of this GPU, measured now: uniformly random 64-byte blocks
        # fetched by lane pairs / 32-byte sectors by single lanes over 8 GiB
        gather = {"gbs_64B_blocks": gpu.measure_gather(8 << 30, 64, 1 << 27, 2), "gbs_32B_sectors": gpu.measure_gather(8 << 30, 32, 1 << 27, 1),
                  "footprint_gib": 8}
    ka_ms = stage["phase_a_kernel"]
    achieved = abytes / (ka_ms / 1000.0) / 1e9
    roofline = {"bound": "hbm", "kernel": "phase_a_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": abytes, "kernel_ms": ka_ms,
                "kernel_share_of_step": ka_ms / ms_per_step, "random_gather_peak": gather}

    # reference's CPU path on a bounded sample of the same workload (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        tmp = tempfile.mkdtemp(prefix="sage2_cpu_")
        s_reads, s_k, s_cfg = sample_workload(15.0, tmp)
        d = reference_run(s_reads, s_k, tmp)
        cpu = {"value": d["input_reads"] / d["t_steps123"], "unit": UNIT, "cores": d.get("threads"), "kind": d["kind"],
               "sample": f"cfg2-shaped sample: {s_cfg['genome_bp']} bp genome, {len(s_reads)} reads, steps 1-3 "
                         f"({d['t_steps123']:.1f} s) without FASTQ parse / text output",
               "breakdown_s": {k2: v for k2, v in d.items() if k2.startswith("t_")}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": dict(cfg, parallelism=(
                           (f"{world} GPU(s): reads replicated, table sharded by key hash, window probes routed to their owners "
                            f"({'kernel stores into peer-memory mailboxes over NVLink' if args.exchange == 'p2p' else 'NCCL all-to-all'}), "
                            f"phase A partitioned by read id ({comm['sent']} B sent per rank and step)") if sharded else
                           "single GPU" if world == 1 else
                           f"{world} GPUs: reads + table replicated, phase A partitioned by read id, "
                           f"one NCCL exchange (all-gather + all-reduce MAX, {comm['sent']} B sent per rank)"),
                       table=args.table,
                       reads_per_step=n_reads, l2_policy="inputs (460 MB ASCII + working set) larger than the 126 MB L2",
                       timing="CUDA events on the library stream around all steps, max over ranks"),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(comm["h2d"]) * world,
                "d2h_bytes_per_step": int(16 * n_edges), "ms_per_step": e2e_dev_ms / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "edges_per_sec": counters["n_edges"] / (ms_per_step / 1000.0),
        "wall_ms_per_step": wall_ms / args.steps,
        "stage_ms": stage,
        "alt_table": alt,
        "per_step_ms": {"device_resident": per_step[0], "e2e": per_step[1]},
        "counters": {kk: counters[kk] for kk in ("good_reads", "unique_reads", "distinct_keys", "compare_calls",
                                                  "window_probes", "n_edges", "left_to_explore", "record_words")},
    }
    print(json.dumps(line), flush=True)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--table", default="replicated", choices=["replicated", "sharded"],
                    help="N > 1: every GPU holds the whole table, or one key-hash shard of it with routed probes (SURVEY 8(e))")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="--table sharded: routed probes through peer-memory mailboxes (kernel stores over NVLink) or NCCL all-to-all")
    ap.add_argument("--no-alt-table", action="store_true", help="N > 1: do not also time the other table layout")
    ap.add_argument("--batch-reads", type=int, default=1 << 20, help="reads per routed batch (--table sharded)")
    ap.add_argument("--no-gather", action="store_true", help="skip the random-gather ceiling microbenchmark")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
