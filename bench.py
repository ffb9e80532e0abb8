#!/usr/bin/env python
"""bench.py — headline measurement of the hehub_b200 hot path (contract: see DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one rank per GPU
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

Step      one forward negacyclic NTT launch over a batch of 4096 single-limb polynomials, N = 4096,
          Q = 576460752272228353 (BASELINE.json configs[1] at the north-star size; bench/ntt_bm.cpp:9-26).
value     NTTs per second, whole job (all ranks), inputs resident in HBM; rotating over 4 slabs of
          128 MiB so every step streams from HBM, not from the 126 MB L2.
e2e       the same metric through the host-buffer C-ABI call (hehub_b200_ntt_host): pinned host words in,
          host words out, H2D + kernel + D2H inside the timed region.
roofline  algorithmic bytes (16*N per transform) / CUDA-event launch time vs MEASURED_PEAKS.json hbm_gbs.
extras    the other BASELINE configs (sweep over N, INTT, ckks mult+relin C3/C5, rescale C4), each timed
          the same way; informational.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q59 = 576460752272228353
LOGN = 12
POLYS = 4096
SLABS = 4
SWEEP_WAVE = 296  # ciphertexts per wave of the config-5 sweep: a multiple of 37 (148 SMs); 23 GiB of workspace of the 180 GB
METRIC = "NTT/s (N=4096, one 59-bit modulus, batch 4096)"
WORKLOAD = "forward negacyclic NTT, N=4096, L=1, Q=576460752272228353, 4096 polynomials per step per GPU"
UNIT = "NTT/s"


def kernel_source_sha16() -> str:
    """Hash of the sources the headline kernel (ntt_fwd_fast_kernel<12, RowsIO<0>>, instantiated in api.cu) is compiled from:
    ties the committed ncu capture of that kernel to the build it was taken from."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "hehub_b200", "csrc")
    for name in ("api.cu", "compat.h", "context.h", "internal.h", "modarith.cuh", "ntt_engine.cuh", "ntt_plan.h"):
        with open(os.path.join(csrc, name), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores (rank 0 only)
# ---------------------------------------------------------------------------------------------
def _cpu_lib():
    from oracle.binding import Oracle, Reference, build_oracle
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhehub_ref.so")) or os.path.isdir("/root/reference/src/fhe"):
        try:
            return Reference(), "reference"
        except Exception:
            pass
    build_oracle()
    return Oracle(), "port"


def _cpu_worker(barrier, rows, steps, warmup, out_q, idx):
    import numpy as np
    lib, _ = _cpu_lib()
    n = 1 << LOGN
    rng = np.random.default_rng(idx)
    x = rng.integers(0, Q59, (rows, n), dtype=np.uint64)
    for _ in range(warmup):
        lib.bench_ntt(LOGN, Q59, x, True)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.bench_ntt(LOGN, Q59, x, True)
    barrier.wait()
    out_q.put(time.perf_counter() - t0)


def cpu_throughput(workers: int, rows: int, steps: int, warmup: int):
    """`workers` independent processes (the reference is single-threaded and not thread-safe:
    global unsynchronised caches, ntt.cpp:107-115), each transforming `rows` polynomials per step."""
    import multiprocessing as mp
    ctxm = mp.get_context("fork")
    barrier = ctxm.Barrier(workers)
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_cpu_worker, args=(barrier, rows, steps, warmup, q, i)) for i in range(workers)]
    for p in procs:
        p.start()
    times = [q.get() for _ in procs]
    for p in procs:
        p.join()
    elapsed = max(times)
    return workers * rows * steps / elapsed, elapsed


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, kind = _cpu_lib()
    workers = os.cpu_count() or 1
    rows = 256  # bounded sample per worker per step (~15 ms of CPU work each)
    value, elapsed = cpu_throughput(workers, rows, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_sample": f"{workers} processes x {rows} polynomials per step (bounded sample of the workload)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind,
                         "sample": f"{workers} independent processes x {rows} transforms per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks: sample NVML while the GPU is under load
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------
def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from hehub_b200.binding import Context, _mod

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (hehub_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    from hehub_b200.binding import bind_host_to_gpu
    numa_cpus = bind_host_to_gpu(local)  # pinned host buffers of the e2e legs next to this GPU's PCIe root
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream()
    ctx = Context(device=local, stream=stream.cuda_stream)
    n = 1 << LOGN
    mod1, mod1p = _mod([Q59])

    def timed(fn, steps, warmup):
        """device time of `steps` calls of fn(i) on the context's stream, max over ranks, in seconds"""
        with torch.cuda.stream(stream):
            for i in range(warmup):
                fn(i)
            torch.cuda.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(steps):
                fn(i)
            e1.record(stream)
            torch.cuda.synchronize()
            barrier()
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    # ---- headline: forward NTT, 4096 polys x N=4096, rotating over SLABS slabs -------------------
    slabs = [torch.empty(POLYS * n, dtype=torch.int64, device="cuda") for _ in range(SLABS)]
    for i, s in enumerate(slabs):
        ctx._call("lcg_fill", n, mod1p, 1, s.data_ptr(), POLYS, 42 + rank * 100003 + i * POLYS, 1)
    ctx.tables_prepare(LOGN, [Q59])
    ctx.synchronize()

    def step(i):
        ctx._call("ntt_fwd_lazy", LOGN, mod1p, 1, slabs[i % SLABS].data_ptr(), POLYS)

    sampler = ClockSampler(local)
    launches0 = ctx.launch_count()
    sampler.start()
    elapsed = timed(step, args.steps, args.warmup)
    launches = ctx.launch_count() - launches0 - args.warmup
    # keep the GPU under the same load a little longer so the clock sampler sees it (untimed)
    t_end = time.perf_counter() + 0.4
    with torch.cuda.stream(stream):
        while time.perf_counter() < t_end:
            for i in range(16):
                step(i)
            torch.cuda.synchronize()
    sampler.stop()
    value = world * POLYS * args.steps / elapsed
    ms_per_step = 1e3 * elapsed / args.steps

    peaks, peak_src = measured_peaks()
    alg_bytes = 16 * n * POLYS  # read + write of every coefficient, SURVEY §8(d)
    achieved = alg_bytes / (elapsed / args.steps) / 1e9
    # DRAM bytes per launch from `ncu --set full` (tools/gpu_round3.sh writes profiles/ntt_fwd_traffic.json together with the
    # hash of the kernel sources it profiled): reported only while it belongs to the sources that are being timed
    traffic, traffic_note = None, "no ncu capture on file"
    try:
        with open(os.path.join(ROOT, "profiles", "ntt_fwd_traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("source_sha16") == kernel_source_sha16():
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("source")
        else:
            traffic_note = "capture on file is of other kernel sources (stale): rerun tools/gpu_round3.sh"
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "kernel": "ntt_fwd_fast_kernel<12, RowsIO<0>>",
                "algorithmic_bytes_per_launch": alg_bytes}
    # the transforms are bound by the integer (FMA-heavy) pipe, not HBM: report that fraction too
    try:
        with open(os.path.join(ROOT, "profiles", "int_peak.json")) as fh:
            ip = json.load(fh)
        roofline["int_pipe_ceiling_ntt_per_s"] = ip["ntt4096_per_s_alu_ceiling"]
        roofline["int_pipe_frac"] = (value / world) / ip["ntt4096_per_s_alu_ceiling"]
        # the same butterfly loop with twiddles and modulus constants in per-thread registers, as a transform holds
        # them (tools/uniform_tw.cu): the ceiling a kernel of this design can actually reach
        with open(os.path.join(ROOT, "profiles", "r1w_uniform_tw.json")) as fh:
            ut = json.load(fh)
        roofline["int_pipe_ceiling_register_operands_ntt_per_s"] = ut["ntt4096_per_s_ceiling"]["all_per_thread"]
        roofline["int_pipe_frac_register_operands"] = (value / world) / ut["ntt4096_per_s_ceiling"]["all_per_thread"]
    except Exception:
        pass

    # ---- e2e: host words in, host words out through hehub_b200_ntt_host ---------------------------
    hx = ctx.pinned((POLYS, n))
    hy = ctx.pinned((POLYS, n))
    rng = np.random.default_rng(rank)
    hx[:] = rng.integers(0, Q59, (POLYS, n), dtype=np.uint64)
    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        ctx.ntt_host(True, LOGN, [Q59], hx, hy)
    torch.cuda.synchronize()
    barrier()
    # three trials of e2e_steps steps each, the median trial is reported: the host link is shared with whatever
    # else the box is doing, and a single disturbed trial would otherwise stand for the path
    trials = []
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.ntt_host(True, LOGN, [Q59], hx, hy)  # returns when hy holds the result
        torch.cuda.synchronize()
        barrier()
        trials.append(max_over_ranks(time.perf_counter() - t0))
    e2e_elapsed = sorted(trials)[1]
    e2e = {"value": world * POLYS * e2e_steps / e2e_elapsed, "unit": UNIT, "h2d_bytes_per_step": POLYS * n * 8,
           "d2h_bytes_per_step": POLYS * n * 8, "steps": e2e_steps, "host_cpus_bound_to_gpu_numa_node": len(numa_cpus),
           "trials_per_s": [world * POLYS * e2e_steps / t for t in trials],
           "call": "hehub_b200_ntt_host (pinned host buffers, chunked H2D | kernel | D2H pipeline on three streams)"}

    # ---- extras: the other BASELINE configs -------------------------------------------------------
    extras = {}
    if args.extras and args.rows:
        extras["rows_c3_shape"] = run_rows(ctx, torch, timed, world, rank, peaks, cpu_ok=(rank == 0 and world == 1 and not args.no_cpu))
    if args.extras:
        extras.update(run_extras(ctx, torch, stream, timed, barrier, world, rank, peaks, dist if world > 1 else None, args.sweep_cts,
                                 cpu_ok=(rank == 0 and world == 1 and not args.no_cpu), parity_ok=not args.no_cpu, local=local))

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        lib, kind = _cpu_lib()
        rows = 1024
        x = np.random.default_rng(7).integers(0, Q59, (rows, n), dtype=np.uint64)
        lib.bench_ntt(LOGN, Q59, x[:8].copy(), True)
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 8.0:
            lib.bench_ntt(LOGN, Q59, x, True)
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": reps * rows / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"{reps * rows} forward NTTs (N=4096) on one host core, {dt:.1f} s; host has {os.cpu_count()} cores"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "l2_policy": f"inputs larger than L2: {SLABS} slabs of {POLYS * n * 8 >> 20} MiB visited round-robin",
                       "sharding": "independent polynomials per rank, no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches) * world,
            "clocks": sampler.summary(), "ct_mult": ct_mult_summary(extras), "extras": extras,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def ct_mult_summary(extras):
    """The ciphertext-multiplication half of BASELINE's metric ("NTT/s and CKKS ct-mult+relin/s ..."), lifted out of `extras`
    so the driver's line carries it at the top level; every figure is whole-job, device-timed like `value` unless it says e2e."""
    out = {}
    c3 = extras.get("c3_ckks_mult_relin_N8192_L4", {})
    c5 = extras.get("c5_ckks_mult_relin_N32768_L12", {})
    sw = extras.get("c5_sweep_N32768_L12", {})
    if "mult_relin" in c3:
        out["c3_mult_relin_per_s"] = c3["mult_relin"]["per_s"]
        out["c3_mult_relin_frac_hbm"] = c3["mult_relin"].get("frac_hbm")
    if "mult_relin_e2e_host_buffers" in c3:
        out["c3_mult_relin_e2e_per_s"] = c3["mult_relin_e2e_host_buffers"]["per_s"]
    if "mult_relin_single_ct" in c3:
        out["c3_mult_relin_one_pair_per_call_us"] = c3["mult_relin_single_ct"]["us_per_call"]
    if "mult_relin" in c5:
        out["c5_mult_relin_per_s"] = c5["mult_relin"]["per_s"]
        out["c5_mult_relin_frac_hbm"] = c5["mult_relin"].get("frac_hbm")
    if "mult_relin_single_ct" in c5:
        out["c5_mult_relin_one_pair_per_call_us"] = c5["mult_relin_single_ct"]["us_per_call"]
    if sw:
        out["c5_sweep_cts"] = sw.get("cts_total")
        out["c5_sweep_mult_relin_per_s"] = sw.get("mult_relin_per_s_device_time")
        out["c5_sweep_checksum_of_checksums"] = sw.get("checksum_of_checksums")
        if "parity_sample" in sw:
            out["c5_sweep_parity_bit_exact"] = sw["parity_sample"]["bit_exact"]
    return out or None


def run_extras(ctx, torch, stream, timed, barrier, world, rank, peaks, dist, sweep_cts=0, cpu_ok=False, parity_ok=False, local=0):
    """Other BASELINE configs, device-resident, same timing discipline; values are whole-job."""
    import numpy as np
    from hehub_b200.binding import _mod, pick_moduli  # moduli come from the product's create_params rule (csrc/params.cu)
    hbm = peaks["hbm_gbs"]
    out = {}
    mod1, mod1p = _mod([Q59])

    # sweep N = 1024 .. 32768, L = 1, 4096 polys (configs[1]); two slabs >= 256 MiB total per size
    sweep = {}
    for logn in (10, 11, 12, 13, 14, 15):
        n = 1 << logn
        nsl = max(2, (256 << 20) // (POLYS * n * 8))
        slabs = [torch.empty(POLYS * n, dtype=torch.int64, device="cuda") for _ in range(nsl)]
        for i, s in enumerate(slabs):
            ctx._call("lcg_fill", n, mod1p, 1, s.data_ptr(), POLYS, 42 + i * POLYS, 1)
        ctx.tables_prepare(logn, [Q59])
        steps = 8 if logn >= 14 else 16
        row = {}
        for name, fwd in (("ntt", True), ("intt", False)):
            if fwd:
                fn = lambda i: ctx._call("ntt_fwd_lazy", logn, mod1p, 1, slabs[i % nsl].data_ptr(), POLYS)
            else:
                fn = lambda i: ctx._call("intt_lazy", logn, mod1p, 1, slabs[i % nsl].data_ptr(), POLYS, 0)
            el = timed(fn, steps, 3)
            per = el / steps
            gbs = 16 * n * POLYS / per / 1e9
            row[name] = {"per_s": world * POLYS / per, "gbs": gbs, "frac_hbm": gbs / hbm, "us_per_launch": per * 1e6}
        sweep[str(n)] = row
        del slabs
    out["ntt_sweep_L1_batch4096"] = sweep

    def ct_bench(tag, logn, bits, pbits, batch, steps, what):
        mods, p = pick_moduli(bits, pbits, ctx.lib)
        mods = [int(m) for m in mods]
        ext = mods + [int(p)]
        L, n = len(mods), 1 << logn
        extm, extp = _mod(ext)
        modm, modp = _mod(mods)
        key = torch.empty(L * 2 * (L + 1) * n, dtype=torch.int64, device="cuda")
        if rank == 0 or dist is None:
            ctx._call("lcg_fill", n, extp, L + 1, key.data_ptr(), L * 2 * (L + 1), 1000, 1)
            ctx.synchronize()
        if dist is not None:  # key-switch key: generated once, broadcast over NCCL / NVLink
            torch.cuda.synchronize()
            dist.broadcast(key, src=0)
            torch.cuda.synchronize()
        ct1 = torch.empty(batch * 2 * L * n, dtype=torch.int64, device="cuda")
        ct2 = torch.empty_like(ct1)
        res = torch.empty_like(ct1)
        ctx._call("lcg_fill", n, modp, L, ct1.data_ptr(), batch * 2 * L, 100 + rank * 7919, 1)
        ctx._call("lcg_fill", n, modp, L, ct2.data_ptr(), batch * 2 * L, 200 + rank * 7919, 1)
        ctx.synchronize()
        r = {}
        if "mult" in what:
            fn = lambda i: ctx._call("ckks_mult_relin", logn, extp, L, ct1.data_ptr(), ct2.data_ptr(), key.data_ptr(), res.data_ptr(), batch)
            el = timed(fn, steps, 2)
            per = el / steps / batch
            gbs = 48 * L * n / per / 1e9
            r["mult_relin"] = {"per_s": world / per, "us_per_ct": per * 1e6, "gbs_algorithmic": gbs, "frac_hbm": gbs / hbm,
                               "transforms_per_ct": L * L + 3 * L + 2}
        if "latency" in what:  # one ciphertext pair per call: the launch-bound end of the same path
            fn = lambda i: ctx._call("ckks_mult_relin", logn, extp, L, ct1.data_ptr(), ct2.data_ptr(), key.data_ptr(), res.data_ptr(), 1)
            el = timed(fn, 50, 5)
            before = ctx.launch_count()
            fn(0)
            ctx.synchronize()
            r["mult_relin_single_ct"] = {"us_per_call": el / 50 * 1e6, "per_s": world * 50 / el,
                                         "kernel_launches_per_call": ctx.launch_count() - before}
        if "rescale" in what:
            fn = lambda i: ctx._call("ckks_rescale", logn, modp, L, ct1.data_ptr(), res.data_ptr(), batch)
            el = timed(fn, steps, 2)
            per = el / steps / batch
            gbs = 16 * n * (2 * L - 1) / per / 1e9
            r["rescale"] = {"per_s": world / per, "us_per_ct": per * 1e6, "gbs_algorithmic": gbs, "frac_hbm": gbs / hbm}
        if "tensor" in what:
            quad = torch.empty(batch * 3 * L * n, dtype=torch.int64, device="cuda")
            fn = lambda i: ctx._call("ckks_tensor", logn, modp, L, ct1.data_ptr(), ct2.data_ptr(), quad.data_ptr(), batch)
            el = timed(fn, steps, 2)
            per = el / steps / batch
            gbs = 56 * L * n / per / 1e9
            r["tensor"] = {"per_s": world / per, "us_per_ct": per * 1e6, "gbs_algorithmic": gbs, "frac_hbm": gbs / hbm}
        if "e2e" in what:
            # the same op through the host-buffer entry point: ciphertexts in pinned host memory, key resident on the device
            h1, h2, ho = (ctx.pinned((batch, 2, L, n)) for _ in range(3))
            rng = np.random.default_rng(5 + rank)
            for k, q in enumerate(mods):
                h1[:, :, k, :] = rng.integers(0, q, (batch, 2, n), dtype=np.uint64)
                h2[:, :, k, :] = rng.integers(0, q, (batch, 2, n), dtype=np.uint64)

            class _Key:  # ckks_mult_relin_host takes a Slab-like object
                ptr = key.data_ptr()
            for _ in range(2):
                ctx.ckks_mult_relin_host(logn, ext, h1, h2, _Key, ho)
            torch.cuda.synchronize()
            barrier()
            reps, t0 = 5, time.perf_counter()
            for _ in range(reps):
                ctx.ckks_mult_relin_host(logn, ext, h1, h2, _Key, ho)
            torch.cuda.synchronize()
            barrier()
            dt = time.perf_counter() - t0
            if dist is not None:
                tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            r["mult_relin_e2e_host_buffers"] = {"per_s": world * batch * reps / dt, "h2d_bytes_per_ct": 2 * 16 * L * n,
                                                "d2h_bytes_per_ct": 16 * L * n, "frac_hbm": 48 * L * n * world * batch * reps / dt / 1e9 / hbm / world}
            if cpu_ok:
                lib, kind = _cpu_lib()
                keyh = key.cpu().numpy().view(np.uint64).reshape(L, 2, L + 1, n)
                lib.ckks_mult_relin(logn, ext, h1[0], h2[0], keyh)
                cr, t0 = 0, time.perf_counter()
                while time.perf_counter() - t0 < 3.0:
                    got = lib.ckks_mult_relin(logn, ext, h1[cr % batch], h2[cr % batch], keyh)
                    cr += 1
                cdt = time.perf_counter() - t0
                ctx.ckks_mult_relin_host(logn, ext, h1, h2, _Key, ho)
                r["cpu_mult_relin"] = {"per_s": cr / cdt, "cores": 1, "kind": kind, "sample": f"{cr} ciphertext pairs, {cdt:.1f} s, one core",
                                       "last_sample_bit_exact_vs_gpu": bool(np.array_equal(got, ho[(cr - 1) % batch]))}
                r["speedup_e2e_vs_one_cpu_core"] = r["mult_relin_e2e_host_buffers"]["per_s"] / (cr / cdt)
                r["speedup_device_resident_vs_one_cpu_core"] = r["mult_relin"]["per_s"] / (cr / cdt)
        if dist is not None and "mult" in what:  # result checksums gathered over NCCL
            chk = res.view(torch.int64)[:: max(1, res.numel() // 4096)].sum().reshape(1)
            allc = [torch.empty_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            r["rank_checksums"] = [int(c.item()) & ((1 << 64) - 1) for c in allc]
        r["shape"] = {"N": n, "L": L, "batch_per_gpu": batch}
        out[tag] = r

    # batch sizes are multiples of the rows one wave of CTAs covers on 148 SMs (N=8192: 2 CTAs per SM, N=16384: one,
    # N=32768: one CTA pair per row), so every transform launch of the composite ops ends on a full wave
    ct_bench("c3_ckks_mult_relin_N8192_L4", 13, [40, 30, 30, 30], 40, 1184, 4, ("mult", "tensor", "e2e", "latency"))
    ct_bench("c4_rescale_N16384_L8", 14, [50] + [40] * 7, 50, 592, 4, ("rescale",))
    ct_bench("c5_ckks_mult_relin_N32768_L12", 15, [50] * 12, 55, 296, 2, ("mult", "tensor", "latency"))

    # config 5 as a sweep: `sweep_cts` independent ciphertext pairs (65 536 in BASELINE; bounded by default so the
    # whole bench stays within minutes) cut into contiguous per-rank ranges, processed in waves, inputs generated
    # on the device; NCCL only broadcasts the key and gathers one checksum per ciphertext (hehub_b200/sweep.py)
    if sweep_cts > 0:
        from hehub_b200.sweep import CtSweep, nccl_collectives
        mods, p = pick_moduli([50] * 12, 55, ctx.lib)
        coll, coll_destroy = None, None
        if dist is not None:
            # the C++ driver talks NCCL itself (csrc/nccl_provider.cpp); the launcher's process group only carries the 128-byte id
            def exchange(raw):
                box = [raw]
                dist.broadcast_object_list(box, src=0)
                return box[0]
            if local == 0:  # the provider library is built in-tree by __graft_entry__.build(); make sure it is there
                import __graft_entry__ as ge
                ge.build_nccl()
            barrier()
            coll, coll_destroy = nccl_collectives(rank, world, torch.cuda.current_device(), exchange)
        sw = CtSweep(ctx, 15, [int(m) for m in mods], int(p), seed=42, rank=rank, world=world, collectives=coll)
        sw.make_key()  # rank 0 generates, ncclBroadcast
        sw.run(min(sweep_cts, SWEEP_WAVE * world), SWEEP_WAVE)  # warm-up (tables, workspaces)
        barrier()
        t0 = time.perf_counter()
        res = sw.run(sweep_cts, SWEEP_WAVE, time_ops=True)  # device time: CUDA events inside the driver, around the mult+relin calls
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([res["op_seconds"], wall], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        op_s, wall_s = float(t[0].item()), float(t[1].item())
        r = {"cts_total": sweep_cts, "scaling": "strong", "wave": SWEEP_WAVE, "mult_relin_per_s_device_time": sweep_cts / op_s,
             "mult_relin_per_s_wall_incl_input_generation_and_checksums": sweep_cts / wall_s,
             "gbs_algorithmic": 48 * 12 * 32768 * sweep_cts / op_s / 1e9,
             "driver": "hehub_b200_sweep_* (C++, csrc/sweep.cu)" + ("; collectives: NCCL from csrc/nccl_provider.cpp" if dist is not None else "")}
        r["frac_hbm_per_gpu"] = r["gbs_algorithmic"] / world / hbm
        if "all_checksums" in res:
            import numpy as np
            r["checksums_gathered"] = int(res["all_checksums"].size)
            r["checksum_of_checksums"] = int(np.bitwise_xor.reduce(res["all_checksums"])) if res["all_checksums"].size else 0
        elif dist is None:
            import numpy as np
            r["checksums_gathered"] = int(res["checksums"].size)
            r["checksum_of_checksums"] = int(np.bitwise_xor.reduce(res["checksums"]))
        if parity_ok and res["count"]:
            # sampled parity at the full config-5 size, on EVERY rank: this rank's first, middle and last
            # ciphertext pairs recomputed by the reference's CPU path and compared through the per-ciphertext checksum; on
            # one GPU the same calls give the CPU side of the config-5 speed-up
            from hehub_b200.sweep import ct_checksum_numpy
            lib, kind = _cpu_lib()
            keyh = sw.key_host()
            picks = sorted({res["first"], res["first"] + res["count"] // 2, res["first"] + res["count"] - 1})
            if world > 1:
                picks = picks[:1] + picks[-1:]  # two per rank keep the N-GPU run short
            okc, t0 = 0, time.perf_counter()
            for idx in picks:
                a, b = sw.one_ct_inputs_host(idx)
                okc += int(ct_checksum_numpy(lib.ckks_mult_relin(15, sw.ext, a, b, keyh)) == int(res["checksums"][idx - res["first"]]))
            cdt = (time.perf_counter() - t0) / len(picks)
            tally = torch.tensor([okc, len(picks)], dtype=torch.int64, device="cuda")
            if dist is not None:
                dist.all_reduce(tally)
            r["parity_sample"] = {"ciphertexts_checked": int(tally[1].item()), "bit_exact": int(tally[0].item()) == int(tally[1].item()),
                                  "ranks_checking": world, "checker": kind}
            if world == 1:
                r["cpu_mult_relin"] = {"per_s": 1.0 / cdt, "cores": 1, "kind": kind,
                                       "sample": f"{len(picks)} ciphertext pairs (incl. regenerating their inputs), one core"}
                r["speedup_device_resident_vs_one_cpu_core"] = r["mult_relin_per_s_device_time"] / (1.0 / cdt)
        sw.close()
        if coll_destroy:
            coll_destroy()
        out["c5_sweep_N32768_L12"] = r
    return out


def run_rows(ctx, torch, timed, world, rank, peaks, cpu_ok):
    """One line per SURVEY §8(a) row: the C-ABI entry point timed on the device (CUDA events, operands
    resident in HBM and larger than L2 for the coefficient-wise rows) and the reference's CPU function
    for the same row timed beside it on one host core (bounded sample).  Shape: BASELINE config C3
    (N = 8192, L = 4, {40,30,30,30}+P40)."""
    import ctypes as C
    import numpy as np
    from hehub_b200.binding import _mod, pick_moduli
    hbm = peaks["hbm_gbs"]
    logn, n = 13, 8192
    mods, p = pick_moduli([40, 30, 30, 30], 40, ctx.lib)
    mods = [int(m) for m in mods]
    ext = mods + [int(p)]
    L = len(mods)
    mm, mp = _mod(mods)
    em, ep = _mod(ext)
    t_plain = 65537
    EW, CT = 512, 256  # polynomials per launch (coefficient-wise rows: 3 x 128 MiB operands), ciphertexts per launch
    u64p = C.POINTER(C.c_uint64)

    def buf(words):
        return torch.empty(words, dtype=torch.int64, device="cuda")

    # two operand sets visited alternately: 2 x 3 x 128 MiB, so a launch never finds its operands in the 126 MB L2
    sets = []
    for s_ in range(2):
        a_, b_, c_ = buf(EW * L * n), buf(EW * L * n), buf(EW * L * n)
        ctx._call("lcg_fill", n, mp, L, a_.data_ptr(), EW * L, 11 + rank + 1000 * s_, 1)
        ctx._call("lcg_fill", n, mp, L, b_.data_ptr(), EW * L, 77 + rank + 1000 * s_, 1)
        ctx._call("lcg_fill", n, mp, L, c_.data_ptr(), EW * L, 99 + rank + 1000 * s_, 1)
        sets.append((a_, b_, c_))
    A = lambda i: sets[i & 1][0].data_ptr()
    B = lambda i: sets[i & 1][1].data_ptr()
    Cc = lambda i: sets[i & 1][2].data_ptr()
    pairs = buf(2 * EW * n)  # (lo, hi) pairs for montgomery128
    ctx._call("lcg_fill", 2 * EW * n, mp, 1, pairs.data_ptr(), 1, 5, 1)
    key = buf(L * 2 * (L + 1) * n)
    ctx._call("lcg_fill", n, ep, L + 1, key.data_ptr(), L * 2 * (L + 1), 1000, 1)
    ct1, ct2, res = buf(CT * 2 * L * n), buf(CT * 2 * L * n), buf(CT * 2 * L * n)
    ctx._call("lcg_fill", n, mp, L, ct1.data_ptr(), CT * 2 * L, 100, 1)
    ctx._call("lcg_fill", n, mp, L, ct2.data_ptr(), CT * 2 * L, 200, 1)
    quad, extout = buf(CT * 3 * L * n), buf(CT * 2 * (L + 1) * n)
    scal = np.array([3, 5, 7, 11], dtype=np.uint64)
    tern = np.random.default_rng(3).integers(-1, 2, n)  # key generation needs a small (ternary) secret, NTT form
    sk_host = ctx.poly_ntt_fwd(logn, mods, np.stack([np.where(tern < 0, q + tern, tern).astype(np.uint64) for q in mods]))
    sk_dev = torch.from_numpy(sk_host.view(np.int64).copy()).cuda()
    ctx._call("ckks_tensor", logn, mp, L, ct1.data_ptr(), ct2.data_ptr(), quad.data_ptr(), CT)
    ctx.synchronize()

    W = EW * L * n  # words per coefficient-wise operand
    gpu = {  # row: (call, units per launch, algorithmic bytes per unit, unit name)
        "a3 ntt_negacyclic_inplace_lazy (poly, L limbs)": (lambda i: ctx._call("ntt_fwd_lazy", logn, mp, L, A(i), EW), EW, 16 * L * n, "poly"),
        "a4 intt_negacyclic_inplace_lazy (poly, L limbs)": (lambda i: ctx._call("intt_lazy", logn, mp, L, A(i), EW, 0), EW, 16 * L * n, "poly"),
        "a5 batched_mul_mod_hybrid_lazy": (lambda i: ctx._call("mulmod_hybrid_lazy", n, mp, L, A(i), B(i), Cc(i), EW), W, 24, "word"),
        "a6 operator+= (lazy add)": (lambda i: ctx._call("add_lazy", n, mp, L, Cc(i), B(i), EW), W, 24, "word"),
        "a6 operator-= (lazy sub)": (lambda i: ctx._call("sub_lazy", n, mp, L, Cc(i), B(i), EW), W, 24, "word"),
        "a6 operator*= (scalar)": (lambda i: ctx._call("mul_scalar_lazy", n, mp, L, Cc(i), scal.ctypes.data_as(u64p), EW), W, 16, "word"),
        "a6 batched_reduce_strict": (lambda i: ctx._call("reduce_strict", n, mp, L, Cc(i), EW), W, 16, "word"),
        "a7 batched_barrett_lazy": (lambda i: ctx._call("barrett_lazy", n, mp, L, Cc(i), EW), W, 16, "word"),
        "a8 batched_montgomery_128_lazy": (lambda i: ctx._call("montgomery128_lazy", mods[0], EW * n, pairs.data_ptr(), Cc(i)), EW * n, 24, "word"),
        "a9 ckks::mult_low_level": (lambda i: ctx._call("ckks_tensor", logn, mp, L, ct1.data_ptr(), ct2.data_ptr(), quad.data_ptr(), CT), CT, 56 * L * n, "ct"),
        "a10 ext_prod_montgomery": (lambda i: ctx._call("ext_prod_montgomery", logn, ep, L, ct1.data_ptr(), key.data_ptr(), extout.data_ptr(), CT), CT, 8 * n * (L + 2 * (L + 1)), "poly"),
        "a11 ckks::rescale_inplace": (lambda i: ctx._call("ckks_rescale", logn, mp, L, ct1.data_ptr(), res.data_ptr(), CT), CT, 16 * n * (2 * L - 1), "ct"),
        "a12 bgv::mod_switch_inplace": (lambda i: ctx._call("bgv_mod_switch", logn, mp, L, t_plain, ct1.data_ptr(), res.data_ptr(), CT), CT, 16 * n * (2 * L - 1), "ct"),
        "a13 ckks::relinearize": (lambda i: ctx._call("ckks_relinearize", logn, ep, L, quad.data_ptr(), key.data_ptr(), res.data_ptr(), CT), CT, 40 * L * n, "ct"),
        "a13 ckks::mult (tensor + relinearize)": (lambda i: ctx._call("ckks_mult_relin", logn, ep, L, ct1.data_ptr(), ct2.data_ptr(), key.data_ptr(), res.data_ptr(), CT), CT, 48 * L * n, "ct"),
        "f1 ckks::rotate": (lambda i: ctx._call("ckks_rotate", logn, ep, L, ct1.data_ptr(), key.data_ptr(), 1, res.data_ptr(), CT), CT, 32 * L * n, "ct"),
        "f2 rns_base_transform (1 -> L moduli)": (lambda i: ctx._call("rns_base_transform_from_single", ext[L], mp, L, B(i), Cc(i), n, EW), EW, 8 * n * (1 + L), "poly"),
        "f2 RlweKsk::RlweKsk (samples supplied)": (lambda i: ctx._call("ksk_generate", logn, ep, L, sk_dev.data_ptr(), sk_dev.data_ptr(), ct1.data_ptr(),
                                                                          ct2.data_ptr(), extout.data_ptr()), 1, 8 * n * (2 * L + 4 * L * (L + 1)), "key"),
        # sk = one polynomial of the key slab; plaintexts / masks / errors = halves of the ciphertext slabs
        "f3 decrypt_core": (lambda i: ctx._call("rlwe_decrypt_core", logn, mp, L, ct1.data_ptr(), key.data_ptr(), res.data_ptr(), CT), CT, 24 * L * n, "ct"),
        "f3 encrypt_core (samples supplied)": (lambda i: ctx._call("rlwe_encrypt_core", logn, mp, L, ct1.data_ptr(), key.data_ptr(), ct2.data_ptr(),
                                                                    ct2.data_ptr() + 8 * CT * L * n, res.data_ptr(), CT), CT, 40 * L * n, "ct"),
    }
    out = {}
    for name, (fn, units, bytes_per_unit, unit) in gpu.items():
        el = timed(fn, 6, 3) / 6
        gbs = units * bytes_per_unit / el / 1e9
        out[name] = {"unit": unit, "gpu_per_s": world * units / el, "us_per_launch": el * 1e6, "units_per_launch": units,
                     "algorithmic_bytes_per_unit": bytes_per_unit, "gbs_algorithmic": gbs, "frac_hbm": gbs / hbm}
    if cpu_ok:
        rows_cpu(out, logn, mods, ext, key.cpu().numpy().view(np.uint64).reshape(L, 2, L + 1, n), t_plain, world)
    return {"shape": {"N": n, "L": L, "moduli_bits": [40, 30, 30, 30], "special_bits": 40},
            "note": "gpu: device-resident operands, CUDA-event time; cpu: the reference's function on one host core, same shape", "rows": out}


def rows_cpu(out, logn, mods, ext, hkey, t_plain, world=1, budget_s=0.4):
    """The reference's own functions for the rows of run_rows, one host core, bounded samples."""
    import ctypes as C
    import numpy as np
    u64p = C.POINTER(C.c_uint64)
    n, L = 1 << logn, len(mods)
    lib, kind = _cpu_lib()
    rng = np.random.default_rng(3)
    ha = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods])   # one polynomial [L][N]
    hb_ = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods])
    hc1 = np.stack([ha, hb_])                                              # one ciphertext [2][L][N]
    hc2 = np.stack([hb_, ha])
    hquad = lib.ckks_tensor(logn, mods, hc1, hc2)
    tern = rng.integers(-1, 2, n)  # a ternary secret in NTT form (key generation needs small coefficients)
    hsk = lib.poly_ntt_fwd(logn, mods, np.stack([np.where(tern < 0, q + tern, tern).astype(np.uint64) for q in mods]))
    hmasks = np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in ext]) for _ in range(L)])
    herrs = np.stack([np.stack([rng.integers(0, 9, n).astype(np.uint64) for _ in ext]) for _ in range(L)])
    hpairs = rng.integers(0, mods[0], 2 * n, dtype=np.uint64)
    word = lambda name, *sig: lib._fn(name, None, *sig)
    f_bar = word("barrett_lazy", C.c_uint64, C.c_size_t, u64p)
    f_str = word("reduce_strict", C.c_uint64, C.c_size_t, u64p)
    f_hyb = word("mul_hybrid_lazy", C.c_uint64, C.c_size_t, u64p, u64p, u64p)
    f_mont = word("montgomery128_lazy", C.c_uint64, C.c_size_t, u64p, u64p)
    wa, wb, wc = ha.copy(), hb_.copy(), np.empty_like(ha)
    ptr = lambda arr, k=None: (arr[k] if k is not None else arr).ctypes.data_as(u64p)
    mod_arr = np.array(mods, dtype=np.uint64)

    def per_limb(f):
        def run():
            for k, q in enumerate(mods):
                f(k, q)
        return run

    cpu = {
        "a3 ntt_negacyclic_inplace_lazy (poly, L limbs)": (lambda: lib.poly_ntt_fwd(logn, mods, ha), 1),
        "a4 intt_negacyclic_inplace_lazy (poly, L limbs)": (lambda: lib.poly_intt(logn, mods, ha), 1),
        "a5 batched_mul_mod_hybrid_lazy": (per_limb(lambda k, q: f_hyb(q, n, ptr(wa, k), ptr(wb, k), ptr(wc, k))), L * n),
        "a7 batched_barrett_lazy": (per_limb(lambda k, q: f_bar(q, n, ptr(wc, k))), L * n),
        "a6 batched_reduce_strict": (per_limb(lambda k, q: f_str(q, n, ptr(wc, k))), L * n),
        "a8 batched_montgomery_128_lazy": (lambda: f_mont(mods[0], n, ptr(hpairs), ptr(wc, 0)), n),
        "a9 ckks::mult_low_level": (lambda: lib.ckks_tensor(logn, mods, hc1, hc2), 1),
        "a10 ext_prod_montgomery": (lambda: lib.ext_prod(logn, ext, ha, hkey), 1),
        "a11 ckks::rescale_inplace": (lambda: lib.ckks_rescale(logn, mods, hc1), 1),
        "a12 bgv::mod_switch_inplace": (lambda: lib.bgv_mod_switch(logn, mods, t_plain, hc1), 1),
        "a13 ckks::relinearize": (lambda: lib.ckks_relinearize(logn, ext, hquad, hkey), 1),
        "a13 ckks::mult (tensor + relinearize)": (lambda: lib.ckks_mult_relin(logn, ext, hc1, hc2, hkey), 1),
        "f1 ckks::rotate": (lambda: lib.ckks_rotate(logn, ext, hc1, hkey, 1), 1),
        "f2 rns_base_transform (1 -> L moduli)": (lambda: lib.base_transform_from_single(ext[L], hpairs[:n], mods), 1),
        "f2 RlweKsk::RlweKsk (samples supplied)": (lambda: lib.ksk_generate(logn, ext, hsk, hsk, hmasks, herrs), 1),
        "f3 decrypt_core": (lambda: lib.rlwe_decrypt_core(logn, mods, hc1, ha), 1),
        "f3 encrypt_core (samples supplied)": (lambda: lib.rlwe_encrypt_core(logn, mods, ha, hb_, hc2[0], hc2[1]), 1),
    }
    if kind == "reference":  # the reference's RnsPolynomial operators (rns.cpp:58-171) through the shim
        f_add = lib._fn("poly_add", C.c_int, C.c_uint, C.c_size_t, u64p, u64p, u64p)
        f_sub = lib._fn("poly_sub", C.c_int, C.c_uint, C.c_size_t, u64p, u64p, u64p)
        f_scl = lib._fn("poly_mul_scalar", C.c_int, C.c_uint, C.c_size_t, u64p, u64p, C.c_uint64)
        cpu["a6 operator+= (lazy add)"] = (lambda: f_add(logn, L, ptr(mod_arr), ptr(wa), ptr(wb)), L * n)
        cpu["a6 operator-= (lazy sub)"] = (lambda: f_sub(logn, L, ptr(mod_arr), ptr(wa), ptr(wb)), L * n)
        cpu["a6 operator*= (scalar)"] = (lambda: f_scl(logn, L, ptr(mod_arr), ptr(wa), 3), L * n)
    for name, (fn, units) in cpu.items():
        fn()
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s:
            fn()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        r = out.setdefault(name, {"unit": "?", "gpu_per_s": float("nan")})
        r["cpu_per_s"] = units / dt
        r["cpu_cores"] = 1
        r["cpu_kind"] = kind
        r["cpu_sample"] = f"{reps} calls of {units} {r['unit']}(s)"
        r["gpu_over_one_core"] = r["gpu_per_s"] / world / r["cpu_per_s"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--extras", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--rows", type=int, default=1, help="0: skip the per-row table (SURVEY 8(a) rows, GPU and CPU side by side)")
    ap.add_argument("--sweep-cts", type=int, default=65536,
                    help="ciphertext pairs in the config-5 sweep extra, whole job (BASELINE: 65536; 0 disables)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
