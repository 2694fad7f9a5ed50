#!/usr/bin/env python
"""bench.py -- reads/sec through the STRling extract hot path (repeat-unit scan) on B200, next to the CPU path.

One "step" = one pass of the scan over the configured synthetic workload (BASELINE.json configs[1]: the
30x-WGS-scale class mix of 150 bp reads), issued as sub-batch launches because a batch addresses bases with 32 bits.
  value     reads/s with all inputs already resident in HBM (CUDA events on the launching stream)
  e2e       the same metric through the host C ABI (pinned host buffers -> H2D -> kernel -> D2H inside the timing)
  roofline  algorithmic bytes (62 B per 150 bp read) / measured launch time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (C restatement of the reference CPU path) on a bounded sample, on this box's cores
`--impl reference` times only that CPU restatement (the Nim reference cannot be built in this image).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_READ = 62  # ceil(150/4)=38 B 2-bit SEQ + 8 B descriptor + 16 B result
READ_LEN = 150
P_CLASSES = [0.8, 0.8 - 0.07, 0.6]
METRIC = "reads/sec through extract+cluster at 150 bp; HBM GB/s vs roofline"
WORKLOAD = ("configs[1]: 30x-WGS-scale synthetic 150 bp reads, config-2 class mix: repeat-unit scan of every read (extract K1) + "
            "clustering of the shard's STR reads (K2-K4) + exchange / all-gather of cluster records")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ CPU arm
_JOB = None  # (flat, off, lens, p) inherited by forked workers (no pickling of the read buffer)


def _oracle_worker(idx):
    from oracle import oracle as orc

    flat, off, lens, p, chunks = _JOB
    c = chunks[idx]
    t0 = time.perf_counter()
    orc.get_repeat_batch(flat, off[c], lens[c], p[c])
    return time.perf_counter() - t0


class OracleTimer:
    """The oracle's get_repeat over the same synthetic class mix, fanned out over `cores` forked processes."""

    def __init__(self, n_reads: int, seed: int, cores: int):
        from oracle import oracle as orc
        from strling_b200 import synth

        orc.build()
        reads, cls, lclip, rclip = synth.make_reads(n_reads, seed=seed)
        stride = 160
        segs, _ = synth.segments_for(reads, lclip, rclip, stride)
        flat, off, lens = synth.segment_ascii(reads, segs, stride)
        p = np.asarray(P_CLASSES)[segs["pclass"]]
        order = np.argsort(segs["base_off"], kind="stable")  # keep each read's segments together
        self.job = (flat, off[order], lens[order], p[order], np.array_split(np.arange(len(off)), cores))
        self.n_reads, self.cores = n_reads, cores

    def run(self, passes: int = 1) -> float:
        global _JOB
        _JOB = self.job
        t0 = time.perf_counter()
        if self.cores == 1:
            for _ in range(passes):
                _oracle_worker(0)
        else:
            with mp.get_context("fork").Pool(self.cores) as pool:
                for _ in range(passes):
                    pool.map(_oracle_worker, range(self.cores))
        return time.perf_counter() - t0


def time_oracle(n_reads: int, seed: int, cores: int, passes: int = 1):
    """Returns (reads/s, seconds, reads processed)."""
    dt = OracleTimer(n_reads, seed, cores).run(passes)
    return n_reads * passes / dt, dt, n_reads * passes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # calibrate so that warmup + steps stay within a couple of minutes
    rate1, _, _ = time_oracle(20_000, seed=1, cores=1)
    per_step = int(min(4_000_000, max(50_000, rate1 * cores * 6.0)))
    timer = OracleTimer(per_step, seed=2, cores=cores)
    times = []
    for i in range(args.warmup + args.steps):
        dt = timer.run()
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = per_step / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "bounded CPU sample of the same read mix (scan only: the path's dominant cost)",
                   "reads_per_step": per_step, "read_len": READ_LEN},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} reads/step of the config-2 mix, oracle (C restatement of utils.nim get_repeat; the Nim reference cannot be built here), {cores} processes"},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None
        self.max_mhz = None

    def _run_nvml(self) -> bool:
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            return False
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for nm, bit in bits.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.005)
        return True

    def _run(self):
        if self._run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(phys_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, before any pinned host buffer is allocated (first touch puts the
    pages on that node): with 8 ranks the end-to-end leg moves ~50 GB/s per GPU through host memory, which must not cross sockets."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(phys_index)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1 and 64 * w + b < n_cpu}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception as e:  # NVML missing, containers without the call, ...: run unbound
        log(f"[bench] NUMA binding skipped: {e}")
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    from strling_b200 import build as sb_build

    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        sb_build.build_lib()  # no-op when strling_b200/libstrgpu.so is up to date (it travels with the repo snapshot)
    import strling_b200 as sb
    from strling_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the scan has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = 0
    if world > 1 and not args.no_numa_bind:
        vis0 = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ok = vis0 and all(x.strip().isdigit() for x in vis0.split(",")) and local < len(vis0.split(","))
        numa_cpus = bind_to_gpu_numa_node(int(vis0.split(",")[local]) if ok else local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    g = sb.StrGpu(local)
    g.set_proportions(P_CLASSES)
    if world > 1:
        g.comm_init_torch()   # the library's own NCCL communicator (strgpu_comm_init): the id travels through torch's store

    # ---- workload: one seeded shard per rank, replicated to the per-GPU read count
    shard_reads = min(args.shard_reads, args.reads_per_gpu)
    n_sub = max(1, args.reads_per_gpu // shard_reads)
    reads_per_gpu = n_sub * shard_reads
    t0 = time.time()
    reads, cls, lclip, rclip = synth.make_reads(shard_reads, seed=2 + rank)
    seq2, nmask, stride = synth.pack_matrix(reads)
    assert nmask is None
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    n_seg = len(segs)
    seq_bytes = shard_reads * stride // 4
    log(f"[rank {rank}] shard: {shard_reads} reads, {n_seg} segments, {seq_bytes / 1e6:.0f} MB packed, x{n_sub} sub-batches "
        f"({time.time() - t0:.1f}s host gen)")
    h_seq = torch.from_numpy(seq2[:seq_bytes].copy()).pin_memory()
    h_segs = torch.from_numpy(segs.view(np.uint8).reshape(-1).copy()).pin_memory()
    h_out = [torch.empty(n_seg * 8, dtype=torch.uint8).pin_memory() for _ in range(3)]
    # device-resident copy of the whole per-GPU workload (every sub-batch has its own HBM region)
    d_seq = torch.empty(n_sub * seq_bytes + 16, dtype=torch.uint8, device=dev)
    d_segs = torch.empty(n_sub * n_seg * 8, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_sub * n_seg * 8, dtype=torch.uint8, device=dev)
    d_shard_seq = h_seq.to(dev)
    d_shard_segs = h_segs.to(dev)
    for b in range(n_sub):
        d_seq[b * seq_bytes:(b + 1) * seq_bytes].copy_(d_shard_seq)
        d_segs[b * n_seg * 8:(b + 1) * n_seg * 8].copy_(d_shard_segs)
    d_seq[n_sub * seq_bytes:].zero_()
    del d_shard_seq, d_shard_segs
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    # ---- cluster leg (configs[2]): the STR reads of this shard (~1 % of the reads) -> candidate loci.  N > 1: the library's
    # sharded path (strgpu_cluster_sharded_device: owner partition, NCCL exchange, per-rank clustering, all-gather of the
    # cluster records, final order), enqueued without a host synchronisation
    n_treads = max(1000, reads_per_gpu // 100)
    t0 = time.time()
    treads = synth.make_treads(max(10, n_treads // 120), seed=40 + rank, noise_reads=n_treads - (n_treads // 120) * 26, unplaced=n_treads // 200)
    n_treads = len(treads)
    h_treads = torch.from_numpy(treads.view(np.uint8).reshape(-1).copy()).pin_memory()
    d_treads = h_treads.to(dev)
    max_treads = n_treads
    if world > 1:
        mt = torch.tensor([n_treads], dtype=torch.int64, device=dev)
        dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        max_treads = int(mt.item())
    # ~1 cluster record per 110 STR reads here; a rank may fill cap / world of it.  The sharded call is a collective: max_n,
    # pair_capacity and cap must be THE SAME on every rank (they size the exchange and gather slots), hence max_treads
    cap_bounds = max(4096, max_treads // 32) * world
    pair_capacity = max_treads // world + max_treads // (2 * world) + 4096   # 50 % slack over an even split (few, large buckets here)
    d_bounds = torch.zeros(cap_bounds * 48, dtype=torch.uint8, device=dev)
    d_nb = torch.zeros(1, dtype=torch.int32, device=dev)
    cparams = sb.StrGpu.cluster_params(window=480, min_support=5, max_clip_dist=190)
    h_bounds = np.zeros(cap_bounds, dtype=sb.BOUNDS_DTYPE)
    t32 = d_treads.view(torch.int32).view(-1, 6)
    log(f"[rank {rank}] cluster leg: {n_treads} treads ({time.time() - t0:.1f}s host gen)")
    cl_events = []
    cl_stats = {}

    def cluster_device_leg():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world > 1:
            g.cluster_sharded_device(d_treads.data_ptr(), n_treads, max_treads, cparams, d_bounds.data_ptr(), cap_bounds, d_nb.data_ptr(),
                                     stream, pair_capacity=pair_capacity)
        else:
            g.cluster_device(d_treads.data_ptr(), n_treads, cparams, d_bounds.data_ptr(), cap_bounds, d_nb.data_ptr(), stream)
        e1.record()
        cl_events.append((e0, e1))

    n_extra = n_seg - shard_reads           # soft-clip segments: the only ones that carry descriptors
    extra_max = int(segs["len"][shard_reads:].max()) if n_extra else 0
    n_extra_dev, extra_max_dev = n_extra, extra_max
    n_treads_strong = max(1000, n_treads // world)   # a 1 / world slice of the STR reads for the strong-scaling pass
    max_strong = max(1000, max_treads // world)      # the same on every rank

    def cluster_strong():
        g.cluster_sharded_device(d_treads.data_ptr(), n_treads_strong, max_strong, cparams, d_bounds.data_ptr(), cap_bounds, d_nb.data_ptr(),
                                 stream, pair_capacity=max_strong // world + max_strong // (2 * world) + 4096)

    # --streams 2: library calls alternate between two contexts on two streams, so that call i's scan kernel (ALU + shared-memory
    # bound, low issue rate) runs next to call i + 1's pre-filter (ALU + popcount bound) on the same SMs (STRGPU_SHARE_SM geometry)
    lanes = [(g, torch.cuda.current_stream())]
    if args.streams > 1:
        for _ in range(args.streams - 1):
            gx = sb.StrGpu(local)
            gx.set_proportions(P_CLASSES)
            lanes.append((gx, torch.cuda.Stream(device=dev)))

    def step_device():
        # the same entry-point family as the end-to-end leg (uniform reads without descriptors + clip descriptors),
        # on device-resident buffers
        main = torch.cuda.current_stream()
        for gx, st in lanes[1:]:
            st.wait_stream(main)
        for b in range(n_sub):
            gx, st = lanes[b % len(lanes)]
            gx.scan_reads_device(d_seq.data_ptr() + b * seq_bytes, shard_reads, READ_LEN, stride, 0, None,
                                 d_segs.data_ptr() + (b * n_seg + shard_reads) * 8, n_extra, extra_max,
                                 d_out.data_ptr() + b * n_seg * 8, st.cuda_stream)
        for gx, st in lanes[1:]:
            main.wait_stream(st)
        cluster_device_leg()

    # end-to-end leg: the descriptor-free uniform-read call (reads packed on a 152-base stride = 38 B/read; only the soft-clip
    # segments carry descriptors), pinned host buffers
    seq2_e, nmask_e, stride_e = synth.pack_matrix(reads, align_bases=4)
    segs_e, _ = synth.segments_for(reads, lclip, rclip, stride_e)
    extra_e = np.ascontiguousarray(segs_e[shard_reads:])
    seq_bytes_e = shard_reads * stride_e // 4
    h_seq_e = torch.from_numpy(seq2_e[:seq_bytes_e + 8].copy()).pin_memory()
    h_extra_e = torch.from_numpy(extra_e.view(np.uint8).reshape(-1).copy()).pin_memory()
    seq_np, extra_np = h_seq_e.numpy(), h_extra_e.numpy().view(sb.SEGMENT_DTYPE)
    extra_max = int(extra_np["len"].max()) if len(extra_np) else 0
    out_np = [o.numpy().view(sb.REPEAT_DTYPE) for o in h_out]

    def step_e2e():
        inflight = []
        for b in range(n_sub):
            if len(inflight) == 3:
                g.scan_wait(inflight.pop(0))
            inflight.append(g.scan_reads_submit(seq_np, shard_reads, READ_LEN, stride_e, 0, None, extra_np, extra_max, out_np[b % 3]))
        for t in inflight:
            g.scan_wait(t)
        # cluster through the host API: H2D of the tread PODs, kernels (N > 1: + the NCCL exchange / all-gather inside the
        # library), D2H of the bounds records
        n_out = ctypes.c_uint32(0)
        if world > 1:
            g._check(g.L.strgpu_cluster_sharded(g.h, h_treads.data_ptr(), n_treads, max_treads, cparams.ctypes.data, h_bounds.ctypes.data,
                                                cap_bounds, ctypes.byref(n_out)))
        else:
            g._check(g.L.strgpu_cluster(g.h, h_treads.data_ptr(), n_treads, cparams.ctypes.data, h_bounds.ctypes.data, cap_bounds,
                                        ctypes.byref(n_out)))
        cl_stats["bounds_e2e"] = int(n_out.value)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, device_events: bool):
        for _ in range(args.warmup):
            fn()
        barrier()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) and local < len(vis.split(",")) else local
        sampler = ClockSampler(phys)
        if rank == 0:
            sampler.start()
        launches0 = sum(gx.launch_count for gx, _ in lanes)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        sec = e0.elapsed_time(e1) / 1e3 if device_events else wall
        launches = sum(gx.launch_count for gx, _ in lanes) - launches0
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        barrier()
        return sec, launches, clocks

    sec_dev, launches, clocks = timed(step_device, True)
    g.device_status(stream)
    torch.cuda.synchronize()
    timed_ev = cl_events[-args.steps:]
    cluster_ms = float(np.mean([a.elapsed_time(b) for a, b in timed_ev]))   # whole cluster leg incl. the collectives (device time)
    if world > 1:
        g.comm_status(stream)          # raises if an exchange / gather slot overflowed
    n_bounds_all = int(d_nb.item())
    if n_bounds_all > cap_bounds:
        raise SystemExit("bench.py: bounds capacity too small")
    cl_stats["bounds_all"] = n_bounds_all
    # N > 1, outside the timing: the sharded result equals ONE GPU clustering the concatenation of every rank's records
    sharded_ok = None
    if world > 1:
        n_all = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(n_all, torch.tensor([n_treads], dtype=torch.int64, device=dev))
        n_max = int(n_all.max())
        pad = torch.zeros((n_max, 6), dtype=torch.int32, device=dev)
        pad[:n_treads] = t32
        everyone = torch.empty((world * n_max, 6), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(everyone, pad)
        if rank == 0:
            cat = torch.cat([everyone[r * n_max: r * n_max + int(n_all[r])] for r in range(world)]).contiguous()
            cap1 = max(1024, cat.shape[0] // 2)
            d_b1 = torch.zeros(cap1 * 48, dtype=torch.uint8, device=dev)
            d_n1 = torch.zeros(1, dtype=torch.int32, device=dev)
            g.cluster_device(cat.data_ptr(), cat.shape[0], cparams, d_b1.data_ptr(), cap1, d_n1.data_ptr(), stream)
            n1 = int(d_n1.item())
            one = d_b1[: n1 * 48].cpu().numpy().view(sb.BOUNDS_DTYPE)
            many = d_bounds[: n_bounds_all * 48].cpu().numpy().view(sb.BOUNDS_DTYPE)
            fields = ("tid", "left", "left_most", "right", "right_most", "center_mass", "n_left", "n_right", "n_total", "repeat", "n_reads")
            sharded_ok = len(one) == len(many) and all(np.array_equal(one[f], many[f]) for f in fields)
            if not sharded_ok:
                raise SystemExit("bench.py: sharded clustering differs from one GPU over the concatenated records")
            del d_b1, cat
        del everyone, pad
    scan_launches = n_sub * args.steps      # library calls; each is a pre-filter kernel + a scan kernel over its survivors
    # spot-check: device-resident results of the last sub-batch equal the host-API results of the same shard
    sec_e2e, _, _ = timed(step_e2e, False)
    last = d_out[(n_sub - 1) * n_seg * 8:].cpu().numpy().view(sb.REPEAT_DTYPE)
    if not np.array_equal(last, out_np[(n_sub - 1) % 3]):
        raise SystemExit("bench.py: device-resident and host-API results differ")

    # parity at bench scale, outside the timing: a random sample of the TIMED shard's device results against the oracle
    parity = None
    if rank == 0 and args.parity_sample > 0:
        from oracle import oracle as orc   # the checker; never on the measured path

        orc.build()
        rng = np.random.default_rng(12345)
        pick = np.sort(rng.choice(n_seg, size=min(args.parity_sample, n_seg), replace=False))
        flat, off, lens = synth.segment_ascii(reads, segs[pick], stride)
        units, counts = orc.get_repeat_batch(flat, off, lens, np.asarray(P_CLASSES)[segs["pclass"][pick]])
        bad = int(((last["unit"][pick] != units) | (last["repeat_count"][pick] != counts)).sum())
        parity = {"n": int(len(pick)), "mismatches": bad, "found": int((counts > 0).sum()),
                  "what": "random segments of the last timed sub-batch (device results) vs the oracle's get_repeat"}
        if bad:
            raise SystemExit(f"bench.py: {bad} of {len(pick)} sampled segments differ from the oracle")

    # strong scaling (configs[2] as written: the SAME 6e8-read job split over the GPUs): reads_per_gpu / world reads per rank
    strong = None
    if world > 1 and not args.no_strong:
        n_sub_strong = max(1, n_sub // world)

        def step_strong():
            main = torch.cuda.current_stream()
            for gx, st in lanes[1:]:
                st.wait_stream(main)
            for b in range(n_sub_strong):
                gx, st = lanes[b % len(lanes)]
                gx.scan_reads_device(d_seq.data_ptr() + b * seq_bytes, shard_reads, READ_LEN, stride, 0, None,
                                     d_segs.data_ptr() + (b * n_seg + shard_reads) * 8, n_extra_dev, extra_max_dev,
                                     d_out.data_ptr() + b * n_seg * 8, st.cuda_stream)
            for gx, st in lanes[1:]:
                main.wait_stream(st)
            cluster_strong()

        sec_strong, _, _ = timed(step_strong, True)
        g.comm_status(stream)
        strong = {"reads_total": n_sub_strong * shard_reads * world, "reads_per_gpu": n_sub_strong * shard_reads,
                  "treads_per_gpu": n_treads_strong, "ms_per_step": 1e3 * sec_strong / args.steps,
                  "value": n_sub_strong * shard_reads * world * args.steps / sec_strong, "unit": "reads/s",
                  "note": "configs[2] as written: the one-GPU job (reads_per_gpu reads) split over the ranks; compare with the N = 1 line's value"}

    joint = None
    if world > 1 and not args.no_joint:
        joint = joint_leg(g, rank, world, local, dev)
    cli = None
    if world == 1 and not args.no_cli:
        cli = cli_leg(local)

    total_reads = reads_per_gpu * world
    value = total_reads * args.steps / sec_dev
    e2e_value = total_reads * args.steps / sec_e2e
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # dominant kernels = the scan (pre-filter + ladder kernel per library call): average time per call = (step time - cluster leg) / calls
    launch_s = (sec_dev - args.steps * cluster_ms / 1e3) / max(1, scan_launches)
    achieved = ALGO_BYTES_PER_READ * shard_reads / launch_s / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        # ncu measured one library call of tj["reads_per_launch"] reads; DRAM traffic is linear in the reads of a call
        traffic = float(tj["dram_bytes_per_launch"]) * shard_reads / float(tj["reads_per_launch"])
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate1, _, _ = time_oracle(20_000, seed=1, cores=1)
        n_cpu = int(min(8_000_000, max(100_000, rate1 * cores * 3.0)))
        passes = 5      # ~15 s of CPU work on every core
        v, dt, n = time_oracle(n_cpu, seed=2, cores=cores, passes=passes)
        cpu = {"value": v, "unit": "reads/s", "cores": cores, "kind": "port", "single_core_value": rate1,
               "sample": f"{n} reads ({passes} passes over {n_cpu} reads of the same config-2 mix) in {dt:.1f}s, oracle (C restatement of the reference CPU path), {cores} processes"}

    line = {
        "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "reads_per_gpu": reads_per_gpu, "segments_per_gpu": n_seg * n_sub, "read_len": READ_LEN,
                   "sub_batches_per_step": n_sub, "reads_per_launch": shard_reads,
                   "l2": f"each launch streams {(seq_bytes + 16 * n_seg) / 1e6:.0f} MB of its own HBM region (> 126 MB L2); {n_sub} distinct regions per step",
                   "streams": f"{len(lanes)} library contexts on {len(lanes)} CUDA streams, calls alternate (call i's ladder kernels overlap call i+1's pre-filter)",
                   "parallelism": f"read batches sharded over {world} GPU(s), no data-path collective",
                   "numa": f"rank pinned to the {numa_cpus} CPUs local to its GPU" if numa_cpus else "unbound"},
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(n_sub * (seq_bytes_e + len(extra_np) * 8) + n_treads * 24),
                "api": "strgpu_scan_reads_submit/wait (3 slots in flight) + " + ("strgpu_cluster_sharded" if world > 1 else "strgpu_cluster") + ", pinned host buffers",
                "d2h_bytes_per_step": int(n_sub * n_seg * 8 + cl_stats.get("bounds_e2e", 0) * 48), "ms_per_step": 1e3 * sec_e2e / args.steps,
                "h2d_gb_per_s_per_gpu": (n_sub * (seq_bytes_e + len(extra_np) * 8) + n_treads * 24) / (sec_e2e / args.steps) / 1e9,
                "bound": "PCIe host-to-device copy of the 2-bit reads (38 B per 150-bp read); kernels and the device-to-host copy of the results overlap it"},
        "cluster": {"treads_per_gpu": n_treads, "ms_per_step": cluster_ms, "treads_per_s": n_treads * world / (cluster_ms / 1e3),
                    "bounds": cl_stats.get("bounds_all"),
                    "api": "strgpu_cluster_sharded_device (owner partition, NCCL exchange of 24-byte STR-read records, per-rank K2-K4, NCCL all-gather of 48-byte cluster records, final order; no host synchronisation)" if world > 1 else "strgpu_cluster_device (enqueued without host synchronisation)",
                    "sharded_equals_single_gpu": sharded_ok},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "repeat_prefilter + ladder_stage<2> + ladder_stage<4> (one library call)",
                     "note": "achieved = 62 B x reads per call / measured time of the call's kernels (CUDA events over the timed region, cluster leg subtracted); integer-pipe bound, not HBM bound: see DESIGN.md"},
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if parity:
        line["parity_sample"] = parity
    if strong:
        line["strong"] = strong
    if joint:
        line["joint"] = joint
    if cli:
        line["cli"] = cli
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# (n_pairs, seed) of `strling debug synth-bam` -> md5 of the .bin the oracle pipeline writes for it (tools/cli_oracle_md5.py)
CLI_BIN_MD5_ORACLE = {(3_000_000, 2): "ab5b3864dcff18faea27cd62b6b20932"}


def cli_leg(local: int, n_pairs: int = 3_000_000):
    """What a user runs: `strling extract` (BGZF inflate + BAM decode + staging on the host cores, scan on the GPU, mate pairing
    replay, .bin) on a synthetic coordinate-sorted BAM in the shape of configs[1] (6x10^6 150-bp reads, `strling debug synth-bam`:
    90 % plain / 7 % messy / 2 % clipped-STR / 1 % STR reads, 1 % of the pairs without coordinates), best of three runs, with
    the binary's own stage report."""
    import re
    import subprocess as sp
    import tempfile

    from strling_b200 import build as sb_build

    cli = sb_build.build_cli()
    d = tempfile.mkdtemp(prefix="bench_cli_")
    bam, out = os.path.join(d, "bench.bam"), os.path.join(d, "bench.bin")
    # deflate level 6: what htslib / samtools write by default (level 1 -- more and shorter matches -- inflates ~25 % slower per read)
    r = sp.run([cli, "debug", "synth-bam", bam, str(n_pairs), "2", "6"], capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit("bench.py: strling debug synth-bam failed: " + r.stderr[-500:])
    best = None
    for _ in range(3):
        r = sp.run([cli, "extract", "-v", "--device", str(local), bam, out], capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("bench.py: strling extract failed: " + r.stderr[-500:])
        perf = json.loads(re.search(r"perf: (\{.*\})", r.stderr).group(1))
        if best is None or perf["scan_pass_s"] < best["scan_pass_s"]:
            best = perf
    res = {"reads": best["reads"], "reads_per_s": best["reads_per_s"], "str_reads": best["str_reads"], "bam_mb": round(os.path.getsize(bam) / 1e6, 1),
           "bam_deflate_level": 6,
           "stages": {k: best[k] for k in ("inflate_s", "stage_s", "submit_s", "gpu_wait_s", "replay_s", "scan_pass_s", "total_s", "threads", "replay_shards")
                      if k in best},
           "command": "strling extract -v <bam> <bin> (best of 3)",
           "bound": "host: BGZF inflate (the repo's own decoder) + BAM decode + staging + sharded mate-pairing replay on the CPU cores; the GPU waits"}
    # the opt-in variant that inflates the BGZF blocks on the GPU: must write the same .bin; reported beside the default, never instead
    # of it (a failure here is recorded, it does not fail the bench line).  Kernel 1 is the one the -m gpu tests cover and
    # `--gpu-inflate` uses by default; kernel 4 (warp-cooperative copies at kernel 1's occupancy) was written after the round's GPU
    # time was spent and has not run on hardware before this line: whatever it does is recorded here.
    ref_bytes = open(out, "rb").read()
    # parity of the whole command line at this size: the md5 of the .bin against the one the ORACLE pipeline gives for the same
    # synthetic BAM (tools/cli_oracle_md5.py: staged segments scanned by oracle/liboracle.so, replayed by `strling debug extract`;
    # computed in the build container, no GPU involved; the records do not depend on the deflate level)
    import hashlib

    res["bin_md5"] = hashlib.md5(ref_bytes).hexdigest()
    expected = CLI_BIN_MD5_ORACLE.get((n_pairs, 2))
    if expected:
        res["bin_md5_oracle"] = expected
        res["bin_equals_oracle"] = res["bin_md5"] == expected
    for key, kernel, extra in (("gpu_inflate", "1", []), ("gpu_inflate_kernel4_untested", "4", ["--batch-reads", "524288"])):
        try:
            out2 = os.path.join(d, f"bench_{key}.bin")
            r = sp.run([cli, "extract", "-v", "--gpu-inflate", *extra, "--device", str(local), bam, out2], capture_output=True, text=True, timeout=180,
                       env=dict(os.environ, STRGPU_INFLATE_KERNEL=kernel))
            if r.returncode != 0:
                res[key] = {"ok": False, "error": r.stderr[-300:]}
            else:
                perf = json.loads(re.search(r"perf: (\{.*\})", r.stderr).group(1))
                res[key] = {"ok": True, "bin_identical": open(out2, "rb").read() == ref_bytes, "reads_per_s": perf["reads_per_s"],
                            "inflate_s": perf["inflate_s"], "scan_pass_s": perf["scan_pass_s"], "kernel": int(kernel),
                            "command": "STRGPU_INFLATE_KERNEL=%s strling extract -v --gpu-inflate %s<bam> <bin> (one run)" % (kernel, " ".join(extra) + (" " if extra else ""))}
            try:
                os.remove(out2)
            except OSError:
                pass
        except Exception as e:  # noqa: BLE001
            res[key] = {"ok": False, "error": repr(e)[:300]}
    for f in (bam, out):
        try:
            os.remove(f)
        except OSError:
            pass
    return res


def joint_leg(g, rank, world, local, dev, n_samples: int = 10, n_pairs: int = 12000):
    """configs[4] in miniature, outside the timing: 10 synthetic samples -> `strling extract` each (this rank's GPU, CLI) ->
    joint merge over all ranks through the library's sharded clustering == `strling merge` on one GPU, line for line."""
    import subprocess as sp
    import tempfile

    import torch.distributed as dist

    from strling_b200 import bamio, joint
    from strling_b200 import build as sb_build

    box = [tempfile.mkdtemp(prefix="bench_joint_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    d = box[0]
    cli = sb_build.build_cli()
    targets = [(f"chr{i + 1}", 1_500_000) for i in range(6)]
    loci = [(i % 6, 100_000 + 61_000 * i, 100_000 + 61_000 * i + 30 + 7 * (i % 9), u)
            for i, u in enumerate(["CAG", "AAAG", "AC", "CCG", "ATTCT", "A", "AAGGG", "CTG", "AAAAT", "GGC", "AT", "CAGG"] * 2)]
    hdr = bamio.sam_header(targets)
    bins = [os.path.join(d, f"s{s}.bin") for s in range(n_samples)]
    t0 = time.time()
    for s in range(rank, n_samples, world):     # every rank prepares its share of the samples
        recs = bamio.simulate_alignments(900 + s, n_pairs, targets, loci, str_pair_frac=0.3, unmapped_pairs=50)
        bam = os.path.join(d, f"s{s}.bam")
        bamio.write_bam(bam, hdr, targets, recs)
        sp.run([cli, "extract", "--device", str(local), bam, bins[s]], check=True, capture_output=True)
    dist.barrier()
    t1 = time.time()
    lines, counts = joint.joint_merge(bins, None, dev, min_support=5, lib=g)
    t2 = time.time()
    out = None
    if rank == 0:
        sp.run([cli, "merge", "-m", "5", "-o", os.path.join(d, "one"), *bins], check=True, capture_output=True)
        one = open(os.path.join(d, "one-bounds.txt")).read().splitlines()[1:]
        same = one == lines
        out = {"samples": n_samples, "pairs_per_sample": n_pairs, "gpus": world, "bounds_lines": len(lines), "equals_single_gpu_merge": same,
               "prepare_s": round(t1 - t0, 2), "joint_merge_s": round(t2 - t1, 3),
               "api": "python -m strling_b200.joint semantics in-process: per-rank .bin blocks -> strgpu_cluster_sharded (merge mode) -> bounds lines"}
        if not same:
            raise SystemExit("bench.py: joint merge over the GPUs differs from `strling merge` on one GPU")
    dist.barrier()
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print to fd 1 (e.g. the NCCL version banner) goes to stderr; stdout carries the one JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-gpu", type=int, default=600_000_000)
    ap.add_argument("--shard-reads", type=int, default=12_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--no-cli", action="store_true", help="skip the `strling extract` command-line leg (N = 1 only)")
    ap.add_argument("--no-joint", action="store_true", help="skip the config-5 joint-merge leg of multi-GPU runs")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling pass (configs[2] as written) of multi-GPU runs")
    ap.add_argument("--parity-sample", type=int, default=200_000, help="segments of the timed shard checked against the oracle outside the timing")
    ap.add_argument("--no-numa-bind", action="store_true")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
