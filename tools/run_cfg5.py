#!/usr/bin/env python
"""BASELINE.json config #5 (the north_star Target): a synthetic human-scale read set on 8 x B200.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/run_cfg5.py [--genome-bp 3100000000] [--coverage 30] [--read-len 150] [-k 70] [--steps 2] [--warmup 1]

3.1 Gbp at 30x is ~620 M reads = 93 GB of characters: no host pipeline holds that, so every rank GENERATES its slice of
the read set on its own GPU (sage2gpu_synth_reads: the genome is a pure function of the seed), packs it
(sage2gpu_pack_slice), and the steps of multi.partitioned_slice_steps follow: all-gather of the packed records, reads
organised by key range, table built by key-hash shard, all-gathers, phase A by id slice, exchange, phases B / C, edge sort.
The packed reads and the table end up replicated on every GPU (north_star: "the packed read set is replicated over NVLink").

No CPU reference can process this input (~180 GB and about an hour), so parity here is by
  * the same generator + the same steps at a size the oracle handles (tests/test_gpu_partitioned.py::
    test_device_generated_slices_equal_oracle) and cfg4 exactness (tests, bench.py parity gate),
  * invariants of an error-free random genome: every rank ends with the same digests; unique reads + duplicates = reads;
    the overlap graph is (nearly) one chain per strand pair: edges ~ unique reads - contigs; reads left for phase C ~ 0.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-bp", type=int, default=3_100_000_000)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("-k", type=int, default=70)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--seed", type=int, default=31)
    ap.add_argument("--table", default="replicated", choices=["replicated", "sharded"],
                    help="replicated: every rank ends with the complete table (all-gather of the shards); sharded: north_star's layout, "
                         "one key-hash shard per GPU, window probes routed through peer-memory mailboxes")
    ap.add_argument("--batch-reads", type=int, default=1 << 18, help="reads per routed batch (--table sharded)")
    ap.add_argument("--low-memory", type=int, default=0, help="1: release the large buffers as soon as no later stage needs them (full size needs it); 2: the workspace too")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from sage2_b200 import api, multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = args.read_len
    n_pairs = int(args.genome_bp * args.coverage / (2 * L))
    chunk = -(-n_pairs // world)
    p0, p1 = min(n_pairs, rank * chunk), min(n_pairs, (rank + 1) * chunk)
    n_slice = 2 * (p1 - p0)
    gpu = api.Sage2Gpu(local)
    if args.low_memory:
        gpu.set_option("low_memory", args.low_memory)
    stream = torch.cuda.ExternalStream(gpu.stream_ptr(), device=dev)
    view = multi.device_view_fn(dev)
    xstats = {}
    t_gen = [0.0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sent = [0]

    def step():
        # the slice's characters exist only until they are packed (93 GB of them over the box at full size): generated at the
        # start of every step (inside the timed region) and released as soon as sage2gpu_pack_slice has returned
        sent[0] = 0
        t0 = time.perf_counter()
        d_bases = torch.empty(max(1, n_slice * L), dtype=torch.uint8, device=dev)
        d_off = torch.empty(n_slice + 1, dtype=torch.int64, device=dev)
        gpu.synth_reads(d_bases.data_ptr(), d_off.data_ptr(), p0, p1 - p0, args.genome_bp, L, 450.0, 30.0, args.seed)
        t_gen[0] = time.perf_counter() - t0
        if args.table == "sharded":
            inner = multi.partitioned_slice_sharded_steps(gpu, rank, world, view, d_bases.data_ptr(), d_off.data_ptr(), n_slice, args.k, True, L,
                                                          batch_reads=args.batch_reads, p2p=True, sent=sent)
        else:
            inner = multi.partitioned_slice_steps(gpu, rank, world, view, d_bases.data_ptr(), d_off.data_ptr(), n_slice, args.k, True, L, sent)

        def steps():
            nonlocal d_bases, d_off
            try:
                req = next(inner)              # sage2gpu_pack_slice has run: the characters can go
            except StopIteration:
                return
            d_bases = d_off = None
            torch.cuda.empty_cache()
            while True:
                val = yield req
                try:
                    req = inner.send(val)
                except StopIteration:
                    return

        if world == 1:
            multi.run_local([steps()])
        else:
            multi.run_dist(steps(), rank, world, dev, xstats)

    # peak device memory of this rank, sampled while the steps run
    import threading
    peak_used, stop = [0], threading.Event()

    def sample():
        torch.cuda.set_device(local)
        while not stop.is_set():
            free_b, total_b = torch.cuda.mem_get_info(dev)
            peak_used[0] = max(peak_used[0], total_b - free_b)
            stop.wait(0.01)
    th = threading.Thread(target=sample, daemon=True)
    th.start()

    for _ in range(args.warmup):
        step()
    xstats.clear()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = api.kernel_launches()
    w0 = time.perf_counter()
    ev0.record(stream)
    stage = {}
    for _ in range(args.steps):
        step()
        for kk, vv in gpu.timers().items():
            stage[kk] = stage.get(kk, 0.0) + vv / args.steps
    ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3 / args.steps
    t = torch.tensor([ev0.elapsed_time(ev1) / args.steps, wall_ms], dtype=torch.float64, device=dev)
    stop.set()
    th.join(timeout=1)
    mem = torch.tensor([float(peak_used[0])], dtype=torch.float64, device=dev)
    c = gpu.counters()
    d = gpu.digest()
    dig = torch.tensor([d["reads"] & 0x7FFFFFFFFFFFFFFF, d["edges"] & 0x7FFFFFFFFFFFFFFF, c["compare_calls"], c["window_probes"], c["fast_path_reads"]],
                       dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
        all_dig = [torch.empty_like(dig) for _ in range(world)]
        dist.all_gather(all_dig, dig)
    else:
        all_dig = [dig]
    launches = api.kernel_launches() - l0
    if rank == 0:
        ms = float(t[0])
        same = all(int(x[0]) == int(all_dig[0][0]) and int(x[1]) == int(all_dig[0][1]) for x in all_dig)
        V = sum(int(x[2]) for x in all_dig)
        probes = sum(int(x[3]) for x in all_dig)
        U, N = c["unique_reads"], 2 * n_pairs
        packed = (L + 3) // 4
        B = 32 * ((packed + 31) // 32)
        abytes = float(U * packed + 32 * probes + B * V + 16 * U)          # SURVEY 8(d), the phase-A launches of all ranks together
        ka = stage["phase_a_kernel"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        line = {
            "workload": f"synthetic {args.genome_bp} bp random genome, {L} bp paired-end, {args.coverage}x, -k {args.k}, generated on the devices",
            "n_gpus": world, "table": args.table, "low_memory": args.low_memory, "input_reads": N, "ms_per_step": ms, "wall_ms_per_step": float(t[1]), "steps": args.steps, "warmup": args.warmup,
            "reads_per_sec": N / (ms / 1e3), "edges_per_sec": c["n_edges"] / (ms / 1e3),
            "generate_s_per_rank_and_step": t_gen[0], "device_bytes_in_use_peak_max_over_ranks": float(mem[0]),
            "nvlink_bytes_contributed_per_rank_and_step": sent[0],
            "stage_ms_rank0": stage, "exchange_wall_ms_per_step_rank0": {kk: vv / args.steps for kk, vv in xstats.items()},
            "gpu_launches_rank0": launches,
            # replicated table: the search kernel does all the phase-A work, its time is the denominator; sharded table: the probes run in
            # the owners' answer kernels, so the whole phase-A stage (routing, answers, barriers, search) is
            "phase_a": {"kernel_ms_rank0": ka, "stage_ms_rank0": stage["phase_a"], "algorithmic_bytes_all_ranks": abytes,
                        "achieved_gbs_aggregate": abytes / ((ka if args.table == "replicated" else stage["phase_a"]) / 1e3) / 1e9,
                        "hbm_copy_peak_gbs_per_gpu": peak,
                        "frac_of_aggregate_copy_peak": abytes / ((ka if args.table == "replicated" else stage["phase_a"]) / 1e3) / 1e9 / (peak * world)},
            "counters_rank0": {kk: c[kk] for kk in ("good_reads", "unique_reads", "distinct_keys", "keys_over_threshold", "n_edges", "left_to_explore",
                                                     "contained_ext", "contained_size", "edges_inserted_c", "transitive_removed", "record_words",
                                                     "phase_c_on_device")},
            "compare_calls_all_ranks": V, "window_probes_all_ranks": probes, "fast_path_reads_all_ranks": sum(int(x[4]) for x in all_dig),
            "invariants": {"all_ranks_same_digests": same, "reads_digest": f"{d['reads']:016x}", "edges_digest": f"{d['edges']:016x}",
                           "good_reads_equal_input": c["good_reads"] == N,
                           "edges_over_unique_reads": c["n_edges"] / max(1, U)},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    gpu.close()


if __name__ == "__main__":
    main()
