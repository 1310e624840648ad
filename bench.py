#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the fused drainage + interference step (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload at N = 1: single bubble, 8192 x 16384 thickness grid (the configuration the metric is quoted on),
f32 field, fused drainage (CPU semantics: gravity + diffusion + pole averaging + TFE fronts) + LUT
interference shading to RGBA8: 12 algorithmic bytes per cell-update.  At N > 1 the sphere grows to
8192*N x 16384 and is cut into latitude bands of 8192 rows per GPU with a halo-row exchange every step
(weak scaling).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

import numpy as np

# This script's stdout carries exactly one JSON line.  C libraries (NCCL's version banner) write to fd 1 too,
# so fd 1 is pointed at stderr for the whole run and the JSON line goes to a saved copy of the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Gcell-updates/s drainage+interference step"
ROWS_PER_GPU = 8192
NUM_PHI = 16384
BYTES_PER_CELL = 12  # read h (4) + write h' (4) + write RGBA8 (4); SURVEY.md section 8d
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("k_march_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log(f"[clocks] NVML unavailable: {e}")

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def pinned(nbytes: int):
    """page-locked host buffer from libsbx (cudaHostAlloc) as a numpy uint8 view."""
    from soap_bubble_simulation_b200 import _lib as L

    p = C.c_void_p()
    L.check(L.lib().sbx_host_alloc(nbytes, C.byref(p)))
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))
    return arr, p


class CpuReference:
    """The reference's CPU path on host cores: DrainageSimulator::step (f64) + per-cell LUT shading of the same
    view, on a bounded sample grid (sample_rows x 16384 of the 8192 x 16384 workload)."""

    def __init__(self, sample_rows: int, threads: int):
        import oracle_lib as O
        from soap_bubble_simulation_b200.drainage import camera_position
        from synth import dt_max, synthetic_field

        self.O, self.nt, self.nphi, self.threads = O, sample_rows, NUM_PHI, threads
        self.ora = O.OracleDrainage(self.nt, self.nphi, threads=threads)
        self.ora.data[:] = synthetic_field(self.nt, self.nphi)
        self.dt = 0.25 * dt_max(self.nt, self.nphi)
        self.lut = O.generate_interference_lut(1.33, 1.0)
        self.cam = camera_position(0.2, 0.0, 0.3)

    def step(self):
        self.ora.step(self.dt, 1)
        return self.O.shade_field(self.ora.data, self.nt, 0, 0.025, self.cam, lut=self.lut, intensity=4.0, mode=1,
                                  threads=self.threads)

    def run(self, steps: int, warmup: int):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        el = time.perf_counter() - t0
        return self.nt * self.nphi * steps / el / 1e9, el / steps

    def calibrated_steps(self, budget_s: float, lo: int = 3, hi: int = 400) -> int:
        t0 = time.perf_counter()
        self.step()
        t = time.perf_counter() - t0
        return max(lo, min(hi, int(budget_s / max(t, 1e-6))))


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample_rows = 2048
    ref = CpuReference(sample_rows, threads)
    # honour --steps/--warmup up to a ~2 minute budget for the whole run
    cap = ref.calibrated_steps(100.0, lo=1, hi=10**6)
    steps = max(1, min(args.steps, cap))
    warmup = max(1, min(args.warmup, 3, max(1, cap // 10)))
    val, sec = ref.run(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gcell/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "single bubble 8192x16384 fused drainage+interference (LUT shading), CPU path "
                               f"timed on a {sample_rows}x{NUM_PHI} sample grid (a quarter of the rows) per step",
                   "grid_per_gpu": [ROWS_PER_GPU, NUM_PHI]},
        "cpu_baseline": {"value": val, "unit": "Gcell/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_rows}x{NUM_PHI} grid, {steps} fused steps (C restatement of "
                                   "DrainageSimulator::step + LUT shading, OpenMP over rows; the Rust "
                                   "reference cannot be built in this image)"},
        "e2e": {"value": val, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="sbx", choices=["sbx", "reference"])
    ap.add_argument("--rgb", default="lut", choices=["lut", "exact3", "none"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="theta rows per GPU")
    ap.add_argument("--phi", type=int, default=NUM_PHI)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo rows: in-kernel peer-memory push (default) or NCCL send/recv per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"[bench] WORLD_SIZE {world} != --gpus {args.gpus}; using WORLD_SIZE")
    warm = max(3, args.warmup)

    import torch
    import torch.distributed as dist

    import soap_bubble_simulation_b200 as S
    from soap_bubble_simulation_b200.sharding import Comm, band_partition
    from synth import dt_max, synthetic_field

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; this package has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    num_theta = args.rows * world
    nphi = args.phi
    rb, re = band_partition(num_theta, world)[rank]
    rows = re - rb
    cells_local = rows * nphi
    cfg = S.SimulationConfig.default()
    sim = S.DrainageSimulator(cfg, precision=S.PREC_F32, device=local_rank, num_theta=num_theta, num_phi=nphi,
                              h_ref=0.0, row_begin=rb, row_end=re)
    dt = 0.25 * dt_max(num_theta, nphi)
    mode = {"lut": S.RGB_LUT, "exact3": S.RGB_EXACT3, "none": S.RGB_NONE}[args.rgb]
    view = S.View(mode=mode, format=S.OUT_RGBA8)
    bytes_per_cell = BYTES_PER_CELL if mode != S.RGB_NONE else 8

    # synthetic band straight into pinned memory (also the e2e input)
    in_arr, in_ptr = pinned(cells_local * 4)
    host_in = in_arr.view(np.float32).reshape(1, rows, nphi)
    host_in[0] = synthetic_field(num_theta, nphi, dtype=np.float32, row0=rb, rows=rows)
    sim.upload(host_in)
    comm = None
    if world > 1:
        comm = Comm(rank, world, local_rank)
        sim.attach_comm(comm)
        if args.halo == "peer":
            comm.enable_peer_halo(sim)

    stream = torch.cuda.Stream()  # a real (non-null) stream so that torch's events see libsbx's launches
    torch.cuda.set_stream(stream)
    sim.set_stream(stream.cuda_stream)

    def run_steps(n):
        if mode == S.RGB_NONE:
            sim.step(dt, n)
        else:
            sim.step_fused(dt, view, n)

    run_steps(warm)
    torch.cuda.synchronize()
    launches0 = sim.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    run_steps(args.steps)
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = sim.launch_count() - launches0
    ms_per_step = ms / args.steps
    total_cells = num_theta * nphi
    value = total_cells * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer entry point: H2D field, one fused step, D2H field + frame
    e2e = None
    if mode != S.RGB_NONE and args.e2e_steps > 0:
        out_h_arr, out_h_ptr = pinned(cells_local * 4)
        out_rgb_arr, out_rgb_ptr = pinned(cells_local * 4)
        out_h = out_h_arr.view(np.float32).reshape(1, rows, nphi)
        out_rgb = out_rgb_arr.reshape(1, rows, nphi, 4)
        sim.step_host(dt, 1, view, host_in, out_h, out_rgb)  # warm
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            sim.step_host(dt, 1, view, host_in, out_h, out_rgb)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": total_cells * args.e2e_steps / el / 1e9, "unit": "Gcell/s",
               "h2d_bytes_per_step": int(cells_local * 4), "d2h_bytes_per_step": int(cells_local * 8),
               "ms_per_step": el / args.e2e_steps * 1e3,
               "what": "sbx_drainage_step_host: pinned host f32 field in, 1 fused step, f32 field + RGBA8 frame out"}
        assert out_rgb[0, rows // 2, :, 3].min() == 255

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    achieved = bytes_per_cell * cells_local / (ms_per_step * 1e-3) / 1e9  # GB/s per GPU
    line = {
        "metric": METRIC, "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"single bubble {num_theta}x{nphi} thickness grid, fused drainage+interference "
                               f"({args.rgb} shading -> RGBA8), latitude bands of {rows} rows per GPU",
                   "grid_per_gpu": [rows, nphi], "semantics": "cpu (DrainageSimulator::step) + TFE fronts",
                   "precision": "f32 storage, h_ref=0", "dt": dt,
                   "l2": f"inputs larger than L2 ({cells_local * 4 >> 20} MiB field per GPU vs 126 MB L2)",
                   "halo_exchange": ("none" if world == 1 else "in-kernel peer-memory push over NVLink"
                                     if args.halo == "peer" else "nccl send/recv per step")},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(), "peak_source": peak_src,
                     "kernel": "k_march_tma<LUT> (one launch per step; pole rows folded in)" if mode == S.RGB_LUT else "k_march_tma",
                     "algorithmic_bytes_per_cell": bytes_per_cell},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        ref = CpuReference(2048, threads)
        cb_steps = ref.calibrated_steps(15.0)  # about 15 s of CPU work
        val, sec = ref.run(cb_steps, 1)
        line["cpu_baseline"] = {"value": val, "unit": "Gcell/s", "cores": threads, "kind": "port",
                                "sample": f"2048x{NUM_PHI} grid (a quarter of the workload's rows), {cb_steps} fused steps, "
                                          f"{sec:.3f} s/step (C restatement of DrainageSimulator::step f64 + LUT shading, "
                                          "OpenMP over rows on all host cores; the Rust reference is single-threaded "
                                          "and cannot be built in this image)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
