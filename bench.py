#!/usr/bin/env python
"""Benchmark of the reduced-order frequency-sweep hot path (BASELINE.json metric: reduced-sweep freq points/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|small]

One "step" = one pass of the four hot stages over one synthetic batch: orthonormalise the snapshot block
(CholeskyQR2 + SVD of R), Galerkin-project Ct/Tt/WP, solve every reduced system, evaluate the S-parameters.
``value`` = sweep points processed by all ranks / device time of K steps (inputs resident in HBM);
``e2e``   = the same through the reference-facing Python call with HOST (pinned) inputs, H2D/D2H inside the timing.
N > 1 (torchrun): weak scaling -- every rank holds one cfg-sized row block of the operators/snapshots and one
block of sweep points; r x r partials are all-reduced, Q halo rows exchanged, S-parameters all-gathered.

``--impl reference`` times the CPU restatement of the reference path (oracle/, numpy/scipy = the very library
calls the pure-Python reference makes) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (grid nx, ny, nz per GPU), r, ports, sweep points per GPU
    "cfg2": dict(grid=(20, 10, 1000), r=64, m=2, f=10000,
                 desc="BASELINE configs[1]: synthetic curl-curl FEM N=200k DOF, r=64 basis, 2 ports, 10k freq points on 1 B200"),
    "cfg3": dict(grid=(25, 20, 250), r=256, m=4, f=12500,
                 desc="BASELINE configs[2]: synthetic N=1M DOF, r=256, 4 ports, 100k freq points sharded across 8 B200 "
                      "(per GPU: 125k rows, 12.5k points; weak scaling below 8 GPUs)"),
    "cfg5": dict(grid=(25, 20, 1000), r=512, m=8, f=125000,
                 desc="BASELINE configs[4]: dense wideband sweep N=4M DOF, r=512, 8 ports, 1M freq points, row-sharded projection "
                      "(per GPU at 8 GPUs: 500k rows, 125k points; weak scaling below 8 GPUs)"),
    "small": dict(grid=(5, 4, 100), r=16, m=2, f=500, desc="smoke-sized workload"),
}
METRIC = "reduced-sweep freq points/sec"
UNIT = "points/s"
FP64_PEAK_TFLOPS = 37.05   # measured on this pool's B200 by tools/fp64_peaks.cu (DMMA m8n8k4 issue loop; gpurun_out/fp64_peaks.jsonl)


def build_inputs(wl, world, rank):
    """Seeded synthetic inputs.  The global problem is `world` copies of the per-GPU grid stacked
    along z; returns the GLOBAL operators (scipy, host), this rank's snapshot rows and the global frequency axis."""
    from morfem_b200 import synthetic, dist as mfd
    nx, ny, nz = wl["grid"]
    ct, tt = synthetic.waveguide_operators(nx, ny, nz * world)
    n = ct.shape[0]
    wp = synthetic.port_matrix(n, wl["m"], 19)
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
    f = synthetic.frequency_points(wl["f"] * world)
    row0, row1 = mfd.even_split(n, world, rank)
    # per-rank seed: rows are independent Gaussian columns, so a per-block seed gives a valid global block
    s_loc = synthetic.snapshot_matrix(row1 - row0, wl["r"], seed=1000 + rank, decay_decades=6.0)
    return in_c, in_gamma, in_b, f, s_loc, n


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}

    def __init__(self, device_index, period=0.02):
        super().__init__(daemon=True)
        self.period = period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.ok = False
        self._stop_evt = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as exc:  # pragma: no cover - no GPU
            self.err = repr(exc)

    def sample(self):
        nv = self.nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in {**self.BAD, **self.NOTE}.items():
            if mask & bit:
                self.reasons.add(name)

    def run(self):
        while self.ok and not self._stop_evt.is_set():
            try:
                self.sample()
            except Exception:
                break
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if self.ok and not self.samples:
            try:
                self.sample()
            except Exception:
                pass
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_hot_path_timed(in_c, in_gamma, in_b, f, s, f_sample):
    """Time the CPU restatement of the reference path (oracle) stage by stage; stages 3+4 on ``f_sample`` points
    spread over the axis and scaled linearly to the full axis (the loop is per-point independent)."""
    from scipy.sparse import csc_array
    from oracle import reference_path as orc
    idx = np.linspace(0, f.size - 1, min(f_sample, f.size)).astype(int)
    fs = f[idx]
    t0 = time.perf_counter()
    q = orc.orthonormal_basis(s)                                                    # implementation.py:226
    t1 = time.perf_counter()
    a0_r, a1_r, a2_r, b_r = orc.galerkin_projection(q, in_c, csc_array(in_c.shape), in_gamma, in_b)   # :181-184
    t2 = time.perf_counter()
    x = orc.reduced_sweep(fs, a0_r, a1_r, a2_r, b_r, lambda t: 1.0, lambda t: t, lambda t: t ** 2, orc.b_coefficient)  # :189-194
    t3 = time.perf_counter()
    gsm = orc.scattering_sweep(fs, x, b_r)                                          # test_helpers.py:60-65
    t4 = time.perf_counter()
    scale = f.size / fs.size
    total = (t2 - t0) + (t4 - t2) * scale
    return {"basis_s": t1 - t0, "projection_s": t2 - t1, "sweep_s_per_point": (t3 - t2) / fs.size,
            "gsm_s_per_point": (t4 - t3) / fs.size, "step_s_full_axis": total, "points_sampled": int(fs.size)}, gsm


def blas_threads():
    try:
        import threadpoolctl
        return max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    in_c, in_gamma, in_b, f, s_loc, n = build_inputs(wl, 1, 0)
    f_sample = min(f.size, args.cpu_points)
    for _ in range(args.warmup):
        cpu_hot_path_timed(in_c, in_gamma, in_b, f, s_loc, max(8, f_sample // 16))
    times = []
    t_start = time.perf_counter()
    for _ in range(args.steps):
        t, _ = cpu_hot_path_timed(in_c, in_gamma, in_b, f, s_loc, f_sample)
        times.append(t)
    wall = time.perf_counter() - t_start
    step_s = float(np.mean([t["step_s_full_axis"] for t in times]))
    value = f.size / step_s
    cores = blas_threads()
    sample = (f"stages 1+2 in full (N={n}, r={wl['r']}), stages 3+4 on {times[0]['points_sampled']} of {f.size} points scaled linearly; "
              f"numpy/scipy = the reference's own library calls (oracle port), real float64 (the reference's dtype)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "note": "CPU arm runs the single-GPU workload on rank 0's host cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "stages": {"basis_ms": float(np.mean([t["basis_s"] for t in times])) * 1e3,
                       "projection_ms": float(np.mean([t["projection_s"] for t in times])) * 1e3,
                       "sweep_us_per_point": float(np.mean([t["sweep_s_per_point"] for t in times])) * 1e6,
                       "gsm_us_per_point": float(np.mean([t["gsm_s_per_point"] for t in times])) * 1e6},
            "wall_s": wall}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- B200 arm
def run_b200(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from scipy.constants import pi, epsilon_0
    from morfem_b200 import device as dv, dist as mfd, _ffi, implementation as impl, test_helpers as th

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _ffi.load()

    in_c, in_gamma, in_b, f, s_loc, n = build_inputs(wl, world, rank)
    from scipy.sparse import csc_array
    cb = impl.coefficient_array(th.b_coefficient, f)
    coeffs = [np.ones_like(f), f, f ** 2, cb, 2 * pi * f * epsilon_0]
    path = mfd.ShardedHotPath([in_c, csc_array(in_c.shape), in_gamma], in_b, n, f.size, coeffs)
    # "c128" (default, the north star's arithmetic): the real synthetic data embedded in complex128, complex128 kernels
    # throughout.  The synthetic operators and snapshots are real, like the reference's data (main.py:21-23), so "auto" /
    # "f64" run the real float64 twins -- what the public API selects by itself for such inputs; the default run times
    # that path too and reports it under "other_dtype".
    real = args.dtype in ("auto", "f64")
    s_dev = dv.real_or_complex_to_device(s_loc, dev, widen=not real)
    f_total = f.size

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    use_graph = world == 1 and not args.no_graph

    def step():
        if use_graph:
            return path.step_graph(s_dev, want_x=False)       # whole step replayed from one CUDA graph
        return path.step_deferred(s_dev, want_x=False, gather=True)   # eager launches, CholeskyQR2 flags verified after the loop

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    launches0 = _ffi.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = _ffi.launch_count() - launches0
    if not use_graph and not path.verify_deferred():
        raise SystemExit("bench: optimistic CholeskyQR2 failed verification on the synthetic snapshot block")
    if use_graph:
        # a graph replay launches the captured kernels without passing through the C ABI's counter
        launches = args.steps * path.launches_per_graph
        if path.verify() is not None:
            raise SystemExit("bench: optimistic CholeskyQR2 failed verification on the synthetic snapshot block")
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = f_total / (ms_step * 1e-3)

    # ---- the same timed loop on the other arithmetic type (same data; complex128 kernels <-> real float64 twins)
    alt = None
    if not args.no_alt_dtype:
        s_alt = dv.real_or_complex_to_device(s_loc, dev, widen=real)

        def step_alt():
            if use_graph:
                return path.step_graph(s_alt, want_x=False)
            return path.step_deferred(s_alt, want_x=False, gather=True)

        for _ in range(max(args.warmup, 3)):
            step_alt()
        barrier()
        a0_, a1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0_.record()
        for _ in range(args.steps):
            step_alt()
        a1_.record()
        barrier()
        ta = torch.tensor([a0_.elapsed_time(a1_)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        ms_alt = float(ta.item()) / args.steps
        if (use_graph and path.verify() is not None) or (not use_graph and not path.verify_deferred()):
            raise SystemExit("bench: optimistic CholeskyQR2 failed verification (alternate dtype)")
        alt = {"dtype": "c128" if real else "f64", "value": f_total / (ms_alt * 1e-3), "unit": UNIT, "ms_per_step": ms_alt,
               "note": "same workload and timed loop on the other arithmetic type: " +
                       ("real data embedded in complex128, complex128 kernels throughout (the north star's kernels)" if real
                        else "real float64 twins of every kernel")}
        del s_alt

    # ---- per-kernel and per-stage timing (CUDA events on the launching stream, separate pass over the same steps)
    prof_steps = max(1, min(args.steps, 5))
    for _ in range(2):                                        # eager warm-up (the timed loop above may have been graph replays)
        path.step(s_dev, want_x=False, gather=True)
    barrier()
    dv.timer = dv.KernelTimer()
    path.stage_events = []
    for _ in range(prof_steps):
        path.step(s_dev, want_x=False, gather=True)           # eager: per-kernel events cannot be recorded inside a replay
    agg = dv.timer.summary()
    dv.timer = None
    # basis and projection overlap (the SVD rotation runs on a side stream under the SpMMs), so they are timed together
    stage_ms = {"basis_plus_projection": 0.0, "sweep": 0.0, "gather": 0.0}
    for ev in path.stage_events:
        for i, key in enumerate(("basis_plus_projection", "sweep", "gather")):
            stage_ms[key] += ev[i].elapsed_time(ev[i + 1]) / prof_steps
    path.stage_events = None
    have_peaks = os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if have_peaks else 6650.0
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if have_peaks else "fallback 6650 GB/s"
    kernels = {}
    dominant, dom_ms, all_ms = None, -1.0, 0.0
    for name, a in agg.items():
        ms_call = a["ms"] / a["calls"]
        kernels[name] = {"calls_per_step": a["calls"] / prof_steps, "ms_per_call": ms_call, "ms_per_step": a["ms"] / prof_steps,
                         "GBps": a["bytes"] / a["calls"] / ms_call / 1e6 if ms_call > 0 else None,
                         "TFLOPs": a["flops"] / a["calls"] / ms_call / 1e9 if ms_call > 0 else None}
        all_ms += a["ms"]
        if a["ms"] > dom_ms:
            dominant, dom_ms = name, a["ms"]
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    except Exception:
        pass
    roof = None
    if dominant is not None:
        a = agg[dominant]
        ms_call = kernels[dominant]["ms_per_call"]
        t_mem = a["bytes"] / a["calls"] / (hbm_peak * 1e9)
        t_flop = a["flops"] / a["calls"] / (FP64_PEAK_TFLOPS * 1e12)
        share = {"share_of_timed_kernels": dom_ms / all_ms, "share_of_step": (dom_ms / prof_steps) / ms_step}
        if t_flop >= t_mem:
            ach = a["flops"] / a["calls"] / (ms_call * 1e-3) / 1e12
            roof = {"kernel": dominant, "bound": "tensor", "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS,
                    "traffic": (traffic.get(dominant) or {}).get("bytes") if (args.workload == "cfg2" and not real) else None, "peak_source": "FP64 tensor-pipe (DMMA) peak measured on this pool's B200 by tools/fp64_peaks.cu; "
                                                    "MEASURED_PEAKS.json holds no FP64 figure and its bf16 figure does not bound an FP64 kernel", **share}
        else:
            ach = a["bytes"] / a["calls"] / (ms_call * 1e-3) / 1e9
            roof = {"kernel": dominant, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": (traffic.get(dominant) or {}).get("bytes") if (args.workload == "cfg2" and not real) else None,
                    "peak_source": peak_src, **share}

    # ---- end to end through the reference-facing call, host (pinned) buffers in, host array out (rank-local job)
    e2e = None
    if world == 1:
        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        def pinned_csc(a):
            a = csc_array(a)
            return csc_array((pin(a.data), pin(a.indices.astype(np.int32)), pin(a.indptr.astype(np.int32))), shape=a.shape)

        s_host = pin(s_loc)
        c_host, g_host, b_host = pinned_csc(in_c), pinned_csc(in_gamma), pinned_csc(in_b)
        out_pinned = torch.empty((f_total, wl["m"], wl["m"]), dtype=torch.complex128).pin_memory()
        k_e2e = max(3, min(args.steps, 20))
        # the call allocates its device buffers on two streams (compute and upload): start from an empty cache and warm the
        # allocator's per-stream pools up, otherwise some timed calls pay a cudaMalloc (seen as 20-50 ms outliers)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        for _ in range(6):
            th.model_order_reduction_gsm_from_snapshots(f, s_host, c_host, g_host, b_host, pinned_out=out_pinned, real_path=real)
        torch.cuda.synchronize()
        import gc
        gc.collect()
        gc.disable()                                           # no collector pause inside the few-millisecond calls
        dv.transfer_bytes["h2d"] = dv.transfer_bytes["d2h"] = 0
        per_call = []
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            tc = time.perf_counter()
            gsm_host = th.model_order_reduction_gsm_from_snapshots(f, s_host, c_host, g_host, b_host, pinned_out=out_pinned, real_path=real)
            per_call.append(time.perf_counter() - tc)           # the call returns after its blocking download
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k_e2e
        e2e = {"value": f_total / dt, "unit": UNIT, "h2d_bytes_per_step": dv.transfer_bytes["h2d"] // k_e2e,
               "d2h_bytes_per_step": dv.transfer_bytes["d2h"] // k_e2e, "ms_per_step": dt * 1e3, "steps": k_e2e,
               "ms_per_call_median": float(np.median(per_call)) * 1e3, "ms_per_call_max": float(np.max(per_call)) * 1e3,
               "ms_per_call": [round(t * 1e3, 3) for t in per_call],
               "call": "morfem_b200.test_helpers.model_order_reduction_gsm_from_snapshots (host ndarrays / scipy csc in pinned memory)"}
        # PCIe floor of that call: nothing but the same pinned host buffers copied to the device (what bounds e2e)
        bufs = [torch.from_numpy(s_host)] + [torch.from_numpy(np.asarray(x)) for a_ in (c_host, g_host, b_host) for x in (a_.data, a_.indices, a_.indptr)]
        dsts = [torch.empty_like(b_, device=dev) for b_ in bufs]
        torch.cuda.synchronize()
        c0_, c1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0_.record()
        for _ in range(3):
            for d_, b_ in zip(dsts, bufs):
                d_.copy_(b_, non_blocking=True)
        c1_.record()
        torch.cuda.synchronize()
        copy_ms = c0_.elapsed_time(c1_) / 3
        copy_bytes = sum(b_.numel() * b_.element_size() for b_ in bufs)
        e2e["h2d_copy_floor"] = {"ms": copy_ms, "bytes": copy_bytes, "GBps": copy_bytes / copy_ms / 1e6,
                                 "note": "the same pinned input buffers copied host->device with nothing else running"}
        del dsts
        # the same call with the FEM operators kept on the device between calls (the model is fixed, the snapshot block is
        # the per-step input): H2D = the snapshot block only
        for _ in range(4):
            th.model_order_reduction_gsm_from_snapshots(f, s_host, c_host, g_host, b_host, pinned_out=out_pinned, real_path=real, operators_resident=True)
        torch.cuda.synchronize()
        dv.transfer_bytes["h2d"] = dv.transfer_bytes["d2h"] = 0
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            th.model_order_reduction_gsm_from_snapshots(f, s_host, c_host, g_host, b_host, pinned_out=out_pinned, real_path=real, operators_resident=True)
        torch.cuda.synchronize()
        dt_res = (time.perf_counter() - t0) / k_e2e
        e2e["operators_resident"] = {"value": f_total / dt_res, "unit": UNIT, "ms_per_step": dt_res * 1e3,
                                     "h2d_bytes_per_step": dv.transfer_bytes["h2d"] // k_e2e, "d2h_bytes_per_step": dv.transfer_bytes["d2h"] // k_e2e,
                                     "note": "same call with operators_resident=True: operators uploaded once, snapshot block uploaded every call"}
        gc.enable()
        # sanity: the e2e result equals the device-resident result
        dev_gsm = out[0].cpu().numpy()
        if not np.allclose(gsm_host, dev_gsm, rtol=1e-9, atol=1e-12):
            raise SystemExit("bench: e2e and device-resident S-parameters disagree")
    else:
        # N > 1: the public multi-GPU call is ShardedHotPath.step on device-resident shards; its e2e adds the per-rank
        # snapshot upload and the S-parameter download
        s_host = torch.from_numpy(np.ascontiguousarray(s_loc)).pin_memory()
        out_pinned = torch.empty((f_total, wl["m"], wl["m"]), dtype=torch.complex128).pin_memory()
        k_e2e = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            sd = s_host.to(dev, non_blocking=False)
            if not real:
                sd = sd.to(torch.complex128)
            gsm_all = path.step(sd, want_x=False, gather=True)[0]
            out_pinned.copy_(gsm_all)
        barrier()
        dt = (time.perf_counter() - t0) / k_e2e
        tt_ = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
        dt = float(tt_.item())
        e2e = {"value": f_total / dt, "unit": UNIT, "h2d_bytes_per_step": int(s_host.numel() * 8), "d2h_bytes_per_step": int(out_pinned.numel() * 16),
               "ms_per_step": dt * 1e3, "steps": k_e2e, "call": "morfem_b200.dist.ShardedHotPath.step (per-rank snapshot rows from pinned host memory)"}

    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        tcpu, _ = cpu_hot_path_timed(in_c, in_gamma, in_b, f, s_loc, min(f.size, args.cpu_points))
        cpu_baseline = {"value": f.size / tcpu["step_s_full_axis"], "unit": UNIT, "cores": blas_threads(), "kind": "port",
                        "sample": f"one pass: stages 1+2 in full (N={n}, r={wl['r']}, svd {tcpu['basis_s']:.2f} s, projection {tcpu['projection_s']:.2f} s), "
                                  f"stages 3+4 on {tcpu['points_sampled']} of {f.size} points ({tcpu['sweep_s_per_point'] * 1e6:.0f} + "
                                  f"{tcpu['gsm_s_per_point'] * 1e6:.0f} us/point) scaled linearly; real float64 like the reference"}

    # ---- stages 1+2 alone, device timed (single rank: replayed from their own CUDA graph; N > 1: eager launches)
    bp_alone_ms = None
    try:
        def step12():
            if use_graph:
                return path.step_graph(s_dev, want_x=False, skip_sweep=True)
            return path.step(s_dev, want_x=False, gather=False, optimistic=True, skip_sweep=True)
        for _ in range(3):
            step12()
        barrier()
        b0_, b1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0_.record()
        for _ in range(args.steps):
            step12()
        b1_.record()
        barrier()
        tb = torch.tensor([b0_.elapsed_time(b1_) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        bp_alone_ms = float(tb.item())
        path._graph = None                                    # the next step_graph call re-captures the full step
    except Exception as exc:                                  # pragma: no cover - reported, not fatal
        bp_alone_ms = None
        print(f"bench: stage-1/2 timing failed: {exc!r}", file=sys.stderr)

    # ---- stages 1+2 inside the timed step against their composite roofline: per kernel
    #      max(algorithmic bytes / measured HBM peak, flops / measured FP64 peak), complex128 operands, real operator values
    sweep_ms_alone = kernels.get("sweep_lu_gsm", {}).get("ms_per_step")
    bp_roof = None
    if sweep_ms_alone:
        n_loc, r_ = n // world, wl["r"]
        wbytes = 16.0 if not real else 8.0
        fl = 1.0 if not real else 0.25                            # real twins: a quarter of the flops
        bw, p64 = hbm_peak * 1e9, FP64_PEAK_TFLOPS * 1e12
        t_roof = max(6 * wbytes * n_loc * r_ / bw, 20.0 * n_loc * r_ * r_ * fl / p64)                       # CholeskyQR2 + rotation
        for a_ in (in_c, in_gamma):
            nnz_loc = a_.nnz / world
            t_roof += max((nnz_loc * 12.0 + 4.0 * (n_loc + 1) + 2 * wbytes * n_loc * r_) / bw, (2.0 if real else 4.0) * nnz_loc * r_ / p64)   # SpMM (real operator values)
            t_roof += max((2 * wbytes * n_loc * r_ + wbytes * r_ * r_) / bw, 8.0 * n_loc * r_ * r_ * fl / p64)     # Q^T (A Q)
        bp_ms = ms_step - sweep_ms_alone - stage_ms["gather"]
        bp_roof = {"ms": bp_alone_ms, "t_roof_ms": t_roof * 1e3, "frac": t_roof * 1e3 / bp_alone_ms if bp_alone_ms else None,
                   "ms_in_step": bp_ms, "frac_in_step": t_roof * 1e3 / bp_ms if bp_ms > 0 else None,
                   "note": "ms: stages 1+2 timed alone on the device (max over ranks; N=1: their own CUDA graph); ms_in_step: whole timed step minus "
                           "the sweep kernel timed alone; roofline = sum over kernels of max(bytes/HBM, flops/FP64) per GPU (SURVEY 8d)"}

    if rank == 0:
        sweep_k = kernels.get("sweep_lu_gsm", {})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "c128" if not real else "f64", "data": "synthetic",
                "config": {"workload": args.workload + ": " + wl["desc"], "N_dof_per_gpu": n // world, "N_dof_total": n, "r": wl["r"], "ports": wl["m"],
                           "freq_points_total": f_total, "step": "basis (CholeskyQR2+SVD) + projection + reduced solves + S-parameters" + (", replayed from one CUDA graph" if use_graph else ""),
                           "l2": "inputs larger than L2 (snapshot block %.0f MB + operators per GPU); no flush" % (s_dev.numel() * s_dev.element_size() / 1e6),
                           "parallelism": "rows of Q/operators and sweep points block-sharded over %d GPU(s)" % world},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu_baseline,
                "other_dtype": alt,
                "basis_plus_projection": bp_roof,
                "stages": {"basis_plus_projection_ms": stage_ms["basis_plus_projection"], "sweep_ms": stage_ms["sweep"],
                           "gather_ms": stage_ms["gather"],
                           "sweep_kernel_points_per_s_per_gpu": (f_total / world) / (sweep_k["ms_per_step"] * 1e-3) if sweep_k.get("ms_per_step") else None},
                "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-points", type=int, default=2000, help="sweep points the CPU arm solves per step (scaled to the full axis)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dtype", default="c128", choices=["auto", "c128", "f64"],
                    help="c128: complex128 kernels throughout (north star); f64: the real float64 twins (the reference's own dtype; "
                         "valid because the synthetic operators and snapshots are real, like the reference's data)")
    ap.add_argument("--no-alt-dtype", action="store_true", help="skip the second timed loop on the other arithmetic type")
    ap.add_argument("--no-graph", action="store_true", help="run the eager (adaptive CholeskyQR) step instead of the CUDA-graph replay at N=1")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if world != args.gpus:
        if args.gpus == 1 and world == 1:
            pass
        else:
            raise SystemExit(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_b200(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
