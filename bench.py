#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the fused drainage + interference step (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workloads (BASELINE.json configs):
  --workload sphere --scaling weak   (default) config 3: single bubble, 8192 rows per GPU x 16384, latitude bands,
                                     halo row exchange every step; at N = 1 the 8192 x 16384 grid the metric is
                                     quoted on
  --workload sphere --scaling strong config 3 as written: the 8192 x 16384 grid cut into N latitude bands
  --workload foam                    config 4: 4096 independent bubbles of 256 x 512, sharded by bubble index
f32 field, fused drainage (CPU semantics: gravity + diffusion + pole averaging + TFE state machine) + LUT
interference shading to RGBA8: 12 algorithmic bytes per cell-update.  One JSON line on stdout (rank 0).

The oracle (oracle/) is used in two places only, never inside a timed region of the CUDA arm: the CPU baseline
(`cpu_baseline`, `--impl reference`) and the post-run `parity_check` of the sharded result.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
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

METRIC = "Gcell-updates/s drainage+interference step"
ROWS_PER_GPU = 8192
NUM_PHI = 16384
FOAM_BUBBLES, FOAM_NT, FOAM_NPHI = 4096, 256, 512
BYTES_PER_CELL = 12  # read h (4) + write h' (4) + write RGBA8 (4); SURVEY.md section 8d
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
KERNEL_SOURCES = ("sbx_march_tma.cu", "sbx_device.cuh", "sbx_tma.cuh", "sbx_internal.cuh", "sbx_launch.h")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def oracle_module():
    """oracle/oracle_lib.py -- checker and CPU baseline only."""
    import importlib.util as u

    spec = u.spec_from_file_location("_sbx_oracle_lib", os.path.join(ROOT, "oracle", "oracle_lib.py"))
    m = u.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def kernel_source_hash() -> str:
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """(dram bytes per launch of the dominant kernel from the committed ncu capture, note).  The capture records the
    hash of the kernel sources it was taken from; a capture of other sources is refused (null + note)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, "no committed ncu capture"
    if t.get("kernel_source_sha16") != kernel_source_hash():
        return None, (f"stale: profiles/traffic.json was captured from kernel sources {t.get('kernel_source_sha16')}, "
                      f"current sources hash to {kernel_source_hash()}")
    return t.get("k_march_dram_bytes_per_launch"), t.get("capture", "profiles/traffic.json")


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
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
    L.check(L.lib().sbx_host_alloc(max(nbytes, 16), C.byref(p)))
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(nbytes, 16),))
    return arr[:nbytes], p


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------ workload description
def workload(args, world: int) -> dict:
    """Grid, sharding and the `config` object, shared by the CUDA arm and the reference arm."""
    from soap_bubble_simulation_b200.synth import dt_max

    if args.workload == "foam":
        nb = args.bubbles
        return {"kind": "foam", "scaling": "strong", "bubbles": nb, "nt": FOAM_NT, "nphi": FOAM_NPHI, "dt": 1e-3,
                "cells_total": nb * FOAM_NT * FOAM_NPHI,
                "config": {"workload": f"foam cluster of {nb} independent bubbles at {FOAM_NT}x{FOAM_NPHI} each, fused "
                                       "drainage+interference (lut shading -> RGBA8), bubbles sharded across the GPUs "
                                       "with no exchange (BASELINE config 4)",
                           "bubbles_per_gpu": nb / world, "grid_per_bubble": [FOAM_NT, FOAM_NPHI],
                           "semantics": "cpu (DrainageSimulator::step): gravity + diffusion + pole averaging + TFE state "
                                        "machine, per bubble h0 = 300..700 nm and R = 12.5..100 mm (SURVEY 8d)",
                           "precision": "f32 storage, offset h_ref = h0 per bubble", "dt": 1e-3,
                           "l2": f"inputs larger than L2 ({nb // world * FOAM_NT * FOAM_NPHI * 4 >> 20} MiB field per GPU "
                                 "vs 126 MB L2)",
                           "halo_exchange": "none (independent bubbles)"}}
    strong = args.scaling == "strong"
    nt = args.rows if strong else args.rows * world
    nphi = args.phi
    dt = 0.25 * dt_max(nt, nphi)
    rows = nt // world
    halo = ("none" if world == 1 else "in-kernel peer-memory read over NVLink" if args.halo == "peer"
            else "nccl send/recv per step")
    return {"kind": "sphere", "scaling": "strong" if strong else "weak", "nt": nt, "nphi": nphi, "dt": dt,
            "cells_total": nt * nphi,
            "config": {"workload": f"single bubble {nt}x{nphi} thickness grid, fused drainage+interference "
                                   f"({args.rgb} shading -> RGBA8), latitude bands of {rows} rows per GPU"
                                   + (" (BASELINE config 3, strong scaling: the grid is fixed)" if strong and world > 1 else ""),
                       "grid_per_gpu": [rows, nphi],
                       "semantics": "cpu (DrainageSimulator::step): gravity + diffusion + pole averaging + TFE state "
                                    "machine; the run starts at t = 0 like a new simulator, so no front is alive yet "
                                    "(first spawn at t = 0.5 s, drainage.rs:285) -- the fronts-alive figure is in "
                                    "other_configs",
                       "precision": "f32 storage, h_ref=0", "dt": dt,
                       "l2": f"inputs larger than L2 ({rows * nphi * 4 >> 20} MiB field per GPU vs 126 MB L2)"
                             if rows * nphi * 4 > 126e6 else
                             f"field of {rows * nphi * 4 >> 20} MiB per GPU is L2-resident (126 MB L2): no flush between steps, "
                             "every step reads what the previous one wrote",
                       "halo_exchange": halo}}


# ------------------------------------------------------------------------------------------------ reference arm
class CpuReference:
    """The reference's CPU path on host cores: DrainageSimulator::step (f64) + per-cell LUT shading of the same
    view, on a bounded sample grid (sample_rows x num_phi of the workload)."""

    def __init__(self, sample_rows: int, nphi: int, threads: int):
        from soap_bubble_simulation_b200.drainage import camera_position
        from soap_bubble_simulation_b200.synth import dt_max, synthetic_field

        O = oracle_module()
        self.O, self.nt, self.nphi, self.threads = O, sample_rows, nphi, threads
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    wl = workload(args, max(world, 1))
    threads = host_threads()
    if wl["kind"] == "foam":
        sample_rows, nphi = 64 * FOAM_NT, FOAM_NPHI  # 64 bubbles' worth of rows
    else:
        sample_rows, nphi = 2048, wl["nphi"]
    ref = CpuReference(sample_rows, nphi, threads)
    # honour --steps/--warmup up to a ~2 minute budget for the whole run
    cap = ref.calibrated_steps(100.0, lo=1, hi=10**6)
    steps = max(1, min(args.steps, cap))
    warmup = max(1, min(args.warmup, max(1, cap // 10)))
    val, sec = ref.run(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gcell/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl["config"],
        "cpu_baseline": {"value": val, "unit": "Gcell/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_rows}x{nphi} sample grid of the workload per step, {steps} fused steps "
                                   "(C restatement of DrainageSimulator::step f64 + per-cell LUT shading, OpenMP over "
                                   "rows on all host cores; the Rust reference is single-threaded and cannot be built in "
                                   "this image)"},
        "e2e": {"value": val, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ CUDA arm helpers
class Ctx:
    """rank / world / torch stream plumbing shared by the measurements."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            log(f"[bench] WORLD_SIZE {self.world} != --gpus {args.gpus}; using WORLD_SIZE")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device; this package has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream()  # a real (non-null) stream so that torch's events see libsbx's launches
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps: int, warm: int, sampler=None) -> float:
        """ms for fn(steps): CUDA events on the launching stream, barrier + synchronize on both sides, max over
        ranks."""
        torch = self.torch
        if warm > 0:
            fn(warm)
        torch.cuda.synchronize()
        self.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        fn(steps)
        e1.record(self.stream)
        torch.cuda.synchronize()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_sphere(S, ctx, nt, nphi, halo, comm_box, precision=None):
    """this rank's latitude band of an nt x nphi sphere, synthetic field uploaded from pinned memory."""
    from soap_bubble_simulation_b200.sharding import Comm, band_partition
    from soap_bubble_simulation_b200.synth import synthetic_field

    rb, re = band_partition(nt, ctx.world)[ctx.rank]
    rows = re - rb
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32 if precision is None else precision,
                              device=ctx.local, num_theta=nt, num_phi=nphi, h_ref=0.0, row_begin=rb, row_end=re)
    in_arr, in_ptr = pinned(rows * nphi * 4)
    host_in = in_arr.view(np.float32).reshape(1, rows, nphi)
    host_in[0] = synthetic_field(nt, nphi, dtype=np.float32, row0=rb, rows=rows)
    sim.upload(host_in)
    if ctx.world > 1:
        if comm_box[0] is None:
            comm_box[0] = Comm(ctx.rank, ctx.world, ctx.local)
        sim.attach_comm(comm_box[0])
        if halo == "peer":
            comm_box[0].enable_peer_halo(sim)
    sim.set_stream(ctx.stream.cuda_stream)
    return sim, host_in, in_ptr, rb, re


def foam_params(b0: int, b1: int, total: int):
    idx = np.arange(b0, b1, dtype=np.float64) * (4095.0 / max(total - 1, 1))
    return (300.0 + 400.0 * idx / 4095.0) * 1e-9, 0.0125 + 0.0875 * idx / 4095.0


def make_foam(S, ctx, total: int):
    from soap_bubble_simulation_b200.sharding import bubble_range

    b0, b1 = bubble_range(total, ctx.world, ctx.rank)
    h0, rad = foam_params(b0, b1, total)
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, device=ctx.local, num_theta=FOAM_NT,
                              num_phi=FOAM_NPHI, h_ref=1.0, batch=b1 - b0, batch_h0_m=h0, batch_radius_m=rad)
    sim.set_stream(ctx.stream.cuda_stream)
    return sim, b0, b1


def peer_status(sim) -> int:
    from soap_bubble_simulation_b200.sharding import Comm

    return Comm.peer_status(sim)


# ------------------------------------------------------------------------------------------------ parity check
EDGE = 8  # owned rows compared on each side of every band


def sphere_edges(field_rows, frame_rows):
    """first / last EDGE owned rows of this rank's band: (field f32, RGBA8)."""
    return (np.array(field_rows[:EDGE]), np.array(field_rows[-EDGE:]),
            np.array(frame_rows[:EDGE]), np.array(frame_rows[-EDGE:]))


def check_sphere(view, nt, nphi, dt, steps, bands, edges, what: str) -> dict:
    """Rank 0: the oracle (window mode: a latitude slab of the global grid around every band edge, stepped `steps`
    times from the same synthetic field) against the first / last EDGE rows every rank holds after `steps` steps."""
    from soap_bubble_simulation_b200.synth import synthetic_field

    O = oracle_module()
    lut = O.generate_interference_lut(1.33, 1.0)
    max_rel, max_lsb, rows_checked = 0.0, 0, 0
    threads = host_threads()
    for (rb, re), (f_top, f_bot, c_top, c_bot) in zip(bands, edges):
        for g0, got_f, got_c in ((rb, f_top, c_top), (re - EDGE, f_bot, c_bot)):
            w0 = max(0, g0 - steps - 1)
            w1 = min(nt, g0 + EDGE + steps + 1)
            ora = O.OracleDrainage(nt, nphi, window=(w0, w1 - w0), threads=threads)
            ora.data[:] = synthetic_field(nt, nphi, dtype=np.float32, row0=w0, rows=w1 - w0).astype(np.float64)
            ora.step(dt, steps)
            ref = np.array(ora.data[g0 - w0:g0 - w0 + EDGE])
            rel = np.abs(got_f.astype(np.float64) / ref - 1.0)
            max_rel = max(max_rel, float(rel.max()))
            px = O.shade_field(ref, nt, g0, 0.025, view.cam_pos, lut=lut, intensity=view.intensity, mode=1, threads=threads)
            max_lsb = max(max_lsb, int(np.abs(px.astype(np.int32) - got_c.astype(np.int32)).max()))
            rows_checked += EDGE
    return {"what": what, "steps": steps, "rows_checked": rows_checked, "max_rel": max_rel, "rgba_max_lsb": max_lsb,
            "tolerance": {"max_rel": 1e-5, "rgba_max_lsb": 1}, "ok": bool(max_rel <= 1e-5 and max_lsb <= 1)}


def gather(ctx, obj):
    if ctx.world == 1:
        return [obj]
    out = [None] * ctx.world
    ctx.dist.all_gather_object(out, obj)
    return out


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_sbx(args):
    ctx = Ctx(args)
    torch = ctx.torch
    import soap_bubble_simulation_b200 as S

    wl = workload(args, ctx.world)
    warm = max(3, args.warmup)
    mode = {"lut": S.RGB_LUT, "exact3": S.RGB_EXACT3, "none": S.RGB_NONE}[args.rgb]
    view = S.View(mode=mode, format=S.OUT_RGBA8)
    bytes_per_cell = BYTES_PER_CELL if mode != S.RGB_NONE else 8
    comm_box = [None]
    dt = wl["dt"]
    total_cells = wl["cells_total"]
    peer = ctx.world > 1 and wl["kind"] == "sphere" and args.halo == "peer"

    if wl["kind"] == "sphere":
        nt, nphi = wl["nt"], wl["nphi"]
        sim, host_in, in_ptr, rb, re = make_sphere(S, ctx, nt, nphi, args.halo, comm_box)
        rows = re - rb
        cells_local = rows * nphi
        batch = 1
    else:
        sim, b0, b1 = make_foam(S, ctx, wl["bubbles"])
        batch = b1 - b0
        rows, nphi, nt = FOAM_NT, FOAM_NPHI, FOAM_NT
        cells_local = batch * rows * nphi
        host_in = None

    def run_steps(n):
        if mode == S.RGB_NONE:
            sim.step(dt, n)
        else:
            sim.step_fused(dt, view, n)

    run_steps(warm)
    torch.cuda.synchronize()
    launches0 = sim.launch_count()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    ms = ctx.timed(run_steps, args.steps, 0, sampler)
    clocks = sampler.stop() if sampler else None
    launches = sim.launch_count() - launches0
    ms_per_step = ms / args.steps
    value = total_cells * args.steps / (ms * 1e-3) / 1e9
    status_timed = peer_status(sim) if peer else 0

    # ---- end to end through the host-buffer entry point: H2D field, one fused step, D2H field + frame
    e2e, e2e_check = None, None
    if mode != S.RGB_NONE and args.e2e_steps > 0:
        if host_in is None:  # foam: the uniform per-bubble start, from pinned memory
            in_arr, in_ptr = pinned(cells_local * 4)
            host_in = in_arr.view(np.float32).reshape(batch, rows, nphi)
            h0, _ = foam_params(b0, b1, wl["bubbles"])
            host_in[:] = h0.astype(np.float32)[:, None, None]
        out_h_arr, out_h_ptr = pinned(cells_local * 4)
        out_rgb_arr, out_rgb_ptr = pinned(cells_local * 4)
        out_h = out_h_arr.view(np.float32).reshape(batch, rows, nphi)
        out_rgb = out_rgb_arr.reshape(batch, rows, nphi, 4)
        sim.set_state(0.0, 0.5, [])  # every e2e step is the first step of a new simulator on the uploaded field
        sim.step_host(dt, 1, view, host_in, out_h, out_rgb)  # warm
        ctx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            sim.set_state(0.0, 0.5, [])
            sim.step_host(dt, 1, view, host_in, out_h, out_rgb)
        torch.cuda.synchronize()
        el = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": total_cells * args.e2e_steps / el / 1e9, "unit": "Gcell/s",
               "h2d_bytes_per_step": int(cells_local * 4), "d2h_bytes_per_step": int(cells_local * 8),
               "ms_per_step": el / args.e2e_steps * 1e3, "steps": args.e2e_steps,
               "what": "sbx_drainage_step_host per rank: pinned host f32 field in, 1 fused step, f32 field + RGBA8 frame "
                       "out, chunk-pipelined H2D / kernels / D2H"
                       + ("; the neighbours' boundary rows of the input are exchanged once per call (NCCL send/recv on "
                          "the upload stream)" if ctx.world > 1 and wl["kind"] == "sphere" else ""),
               "bytes_per_step_are": "per GPU"}
        assert out_rgb[0, rows // 2, :, 3].min() == 255
        if wl["kind"] == "sphere" and not args.no_parity_check:
            edges = gather(ctx, sphere_edges(out_h[0], out_rgb[0]))
            bands = gather(ctx, (rb, re))
            if ctx.rank == 0:
                e2e_check = check_sphere(view, nt, nphi, dt, 1, bands, edges,
                                         "host-buffer step (the e2e path): output rows at every band edge vs the oracle")

    # ---- parity of the device-resident sharded path (the timed path): same field, P steps, every band edge
    parity = None
    if wl["kind"] == "sphere" and mode == S.RGB_LUT and not args.no_parity_check:
        P = min(warm + args.steps, 24)
        sim.upload(host_in)
        sim.set_state(0.0, 0.5, [])
        sim.step_fused(dt, view, P)
        got = sim.download(np.float32)[0]
        px = sim.download_rgb()[0]
        st = peer_status(sim) if peer else 0
        edges = gather(ctx, sphere_edges(got, px))
        bands = gather(ctx, (rb, re))
        stats = gather(ctx, (st, status_timed))
        if ctx.rank == 0:
            parity = check_sphere(view, nt, nphi, dt, P, bands, edges,
                                  "device-resident sharded steps (the timed path): first/last rows of every band vs the "
                                  "oracle run on the matching slab of the same synthetic field")
            parity["peer_status"] = {"after_check": [s[0] for s in stats], "after_timed_region": [s[1] for s in stats]}
            parity["ok"] = bool(parity["ok"] and not any(s[0] or s[1] for s in stats))
            if e2e_check is not None:
                parity["e2e"] = e2e_check
                parity["ok"] = bool(parity["ok"] and e2e_check["ok"])
    elif wl["kind"] == "foam" and mode == S.RGB_LUT and not args.no_parity_check:
        # a few bubbles of every rank against the whole-sphere oracle (256 x 512 finishes in milliseconds)
        P = min(warm + args.steps, 24)
        sim.close()
        sim, b0, b1 = make_foam(S, ctx, wl["bubbles"])  # back to the uniform per-bubble start
        sim.step_fused(dt, view, P)
        got = sim.download(np.float64)
        px = sim.download_rgb()
        pick = sorted(set([0, batch // 2, batch - 1]))
        samples = gather(ctx, [(b0 + k, np.array(got[k]), np.array(px[k])) for k in pick])
        if ctx.rank == 0:
            O = oracle_module()
            lut = O.generate_interference_lut(1.33, 1.0)
            max_rel, max_inc, max_lsb, nb = 0.0, 0.0, 0, 0
            for per_rank in samples:
                for gb, f, c in per_rank:
                    h0, rad = foam_params(gb, gb + 1, wl["bubbles"])
                    ora = O.OracleDrainage(FOAM_NT, FOAM_NPHI, h0_m=float(h0[0]), radius_m=float(rad[0]))
                    ora.step(dt, P)
                    max_rel = max(max_rel, float(np.abs(f / ora.data - 1).max()))
                    dref = ora.data - h0[0]
                    max_inc = max(max_inc, float(np.abs((f - h0[0]) - dref).max() / np.abs(dref).max()))
                    ref = O.shade_field(ora.data, FOAM_NT, 0, float(rad[0]), view.cam_pos, lut=lut,
                                        intensity=view.intensity, mode=1)
                    max_lsb = max(max_lsb, int(np.abs(ref.astype(np.int32) - c.astype(np.int32)).max()))
                    nb += 1
            parity = {"what": "foam: first / middle / last bubble of every rank vs the whole-sphere oracle", "steps": P,
                      "bubbles_checked": nb, "max_rel": max_rel, "increment_rel_err": max_inc, "rgba_max_lsb": max_lsb,
                      "tolerance": {"max_rel": 1e-5, "increment_rel_err": 1e-3, "rgba_max_lsb": 1},
                      "ok": bool(max_rel <= 1e-5 and max_inc <= 1e-3 and max_lsb <= 1)}

    sim.close()
    others = None
    if args.other_configs and wl["kind"] == "sphere" and wl["scaling"] == "weak" and mode == S.RGB_LUT:
        others = other_configs(ctx, S, args, view, comm_box)

    if ctx.rank != 0:
        if comm_box[0] is not None:
            comm_box[0].close()
        ctx.close()
        return

    peak, peak_src = hbm_peak()
    achieved = bytes_per_cell * cells_local / (ms_per_step * 1e-3) / 1e9  # GB/s per GPU
    traffic, traffic_note = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "Gcell/s", "n_gpus": ctx.world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl["config"],
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                     "kernel": "k_march_tma<float,LUT> (one launch per step; pole rows folded in)" if mode == S.RGB_LUT else "k_march_tma",
                     "algorithmic_bytes_per_cell": bytes_per_cell, "per": "GPU"},
    }
    if parity is not None:
        line["parity_check"] = parity
    if others is not None:
        line["other_configs"] = others
    if ctx.world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        if wl["kind"] == "foam":
            ref = CpuReference(64 * FOAM_NT, FOAM_NPHI, threads)
        else:
            ref = CpuReference(2048, nphi, threads)
        cb_steps = ref.calibrated_steps(15.0)  # about 15 s of CPU work
        val, sec = ref.run(cb_steps, 1)
        line["cpu_baseline"] = {"value": val, "unit": "Gcell/s", "cores": threads, "kind": "port",
                                "sample": f"{ref.nt}x{ref.nphi} sample grid of the workload, {cb_steps} fused steps, "
                                          f"{sec:.3f} s/step (C restatement of DrainageSimulator::step f64 + LUT shading, "
                                          "OpenMP over rows on all host cores; the Rust reference is single-threaded "
                                          "and cannot be built in this image)"}
    emit(line)
    bad = parity is not None and not parity["ok"]
    if bad:
        log(f"[bench] PARITY CHECK FAILED: {json.dumps(parity)}")
    if comm_box[0] is not None:
        comm_box[0].close()
    ctx.close()
    if bad:
        sys.exit(3)


def other_configs(ctx, S, args, view, comm_box):
    """Short device-resident measurements of the other BASELINE configs next to the headline, so that the driver's run
    carries them too.  Same timing rules (CUDA events, warm-up, max over ranks); a few seconds in total."""
    from soap_bubble_simulation_b200.synth import dt_max, synthetic_field

    peak, _ = hbm_peak()
    out = []

    def rec(name, cells_per_gpu, total_cells, bpc, ms_per_step, note=""):
        gbs = bpc * cells_per_gpu / (ms_per_step * 1e-3) / 1e9
        out.append({"config": name, "ms_per_step": ms_per_step, "gcell_per_s": total_cells / (ms_per_step * 1e-3) / 1e9,
                    "bytes_per_cell": bpc, "achieved_gbs_per_gpu": gbs, "frac_of_measured_hbm": gbs / peak, "note": note})

    if ctx.world > 1:
        # config 3 as written: the 8192 x 16384 grid, strong scaling over the ranks
        nt, nphi = ROWS_PER_GPU, NUM_PHI
        sim, _, _, rb, re = make_sphere(S, ctx, nt, nphi, args.halo, comm_box)
        dt = 0.25 * dt_max(nt, nphi)
        ms = ctx.timed(lambda n: sim.step_fused(dt, view, n), 200, 20) / 200
        st = peer_status(sim) if args.halo == "peer" else 0
        rec(f"config 3 strong scaling: 8192x16384 over {ctx.world} GPUs ({re - rb} rows per GPU), fused LUT -> RGBA8",
            (re - rb) * nphi, nt * nphi, 12, ms, f"halo: {args.halo}; peer_status {st}")
        sim.close()
        # config 4: 4096 bubbles over the ranks
        sim, b0, b1 = make_foam(S, ctx, FOAM_BUBBLES)
        ms = ctx.timed(lambda n: sim.step_fused(1e-3, view, n), 40, 5) / 40
        rec(f"config 4 foam: 4096 bubbles x 256x512 over {ctx.world} GPUs ({b1 - b0} per GPU), fused LUT -> RGBA8",
            (b1 - b0) * FOAM_NT * FOAM_NPHI, FOAM_BUBBLES * FOAM_NT * FOAM_NPHI, 12, ms, "independent bubbles, no exchange")
        sim.close()
        return out

    def sphere(name, nt, nphi, steps, bpc, prec=S.PREC_F32, rgb=True, note="", state=None):
        sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, num_theta=nt, num_phi=nphi, h_ref=0.0)
        sim.upload(synthetic_field(nt, nphi, dtype=np.float32))
        sim.set_stream(ctx.stream.cuda_stream)
        if state is not None:
            sim.set_state(*state)
        dt = min(1e-3, 0.25 * dt_max(nt, nphi))
        fn = (lambda n: sim.step_fused(dt, view, n)) if rgb else (lambda n: sim.step(dt, n))
        ms = ctx.timed(fn, steps, 5) / steps
        extra = f"; {sim.active_tfe_count()} fronts alive" if state is not None else ""
        rec(name, nt * nphi, nt * nphi, bpc, ms, note + extra)
        sim.close()

    # the TFE state a default-config simulator has at t = 15 s (fronts alive), taken from a small resident run
    small = S.DrainageSimulator(S.SimulationConfig.default(), num_theta=32, num_phi=64)
    small.step(5e-3, 3000)
    state = small.get_state()
    small.close()
    sphere("config 3 with marginal-regeneration fronts alive (TFE state of t = 15 s restored): 8192x16384 fused LUT -> RGBA8",
           8192, 16384, 50, 12, state=state, note="the k_march_tma<TFE=true> instantiation")
    sphere("config 3 drainage only (no colour stage): 8192x16384", 8192, 16384, 50, 8, rgb=False)
    sphere("config 3 f64 storage: 8192x16384 fused LUT -> RGBA8", 8192, 16384, 30, 20, prec=S.PREC_F64)
    sphere("config 2: 2048x4096 fused LUT -> RGBA8", 2048, 4096, 400, 12,
           note="64 MiB working set is L2-resident: the HBM fraction is nominal")
    # config 1: config/default.json, 1000 steps per call + one frame
    sim = S.DrainageSimulator(S.SimulationConfig.default())
    sim.set_stream(ctx.stream.cuda_stream)
    ms = ctx.timed(lambda n: sim.step(1e-3, n), 1000, 50) / 1000
    rec("config 1: 128x256 config/default.json, 1000 drainage steps per call (TFE on, f32 offset storage)", 128 * 256,
        128 * 256, 8, ms, "field resident in shared memory (sbx_resident.cu): exchange-latency-bound, no HBM traffic inside the call")

    def frames(n):
        for _ in range(n):
            sim.step_fused(1e-3, view, 10)

    ms = ctx.timed(frames, 100, 10) / 1000
    rec("config 1: 10 steps per frame + LUT frame (GPUDrainageSimulator::step's steps_per_frame loop), per step",
        128 * 256, 128 * 256, 12, ms, "launch-bound")
    sim.close()
    # config 4: one GPU's share (512 bubbles) and the whole cluster on one GPU
    for B in (512, 4096):
        h0, rad = foam_params(0, B, B)
        sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=FOAM_NT, num_phi=FOAM_NPHI,
                                  h_ref=1.0, batch=B, batch_h0_m=h0, batch_radius_m=rad)
        sim.set_stream(ctx.stream.cuda_stream)
        ms = ctx.timed(lambda n: sim.step_fused(1e-3, view, n), 20, 3) / 20
        rec(f"config 4 foam: {B} bubbles x 256x512 on one GPU, fused LUT -> RGBA8", B * FOAM_NT * FOAM_NPHI,
            B * FOAM_NT * FOAM_NPHI, 12, ms, "one GPU's share at 8 GPUs" if B == 512 else "the whole cluster")
        sim.close()
    # config 5: 64-wavelength LUT build + LUT shading of the 8192 x 16384 field per view
    wl64 = 380.0 + 400.0 * np.arange(64) / 63.0
    S.generate_spectral_lut(1.33, 1.0, wl64, 256, 64, 2000.0)
    t0 = time.perf_counter()
    for _ in range(10):
        lut = S.generate_spectral_lut(1.33, 1.0, wl64, 256, 64, 2000.0)
    t_lut = (time.perf_counter() - t0) / 10
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=8192, num_phi=16384, h_ref=0.0)
    sim.upload(synthetic_field(8192, 16384, dtype=np.float32))
    sim.set_lut(lut)
    sim.set_stream(ctx.stream.cuda_stream)
    views = [S.View(cam_pos=S.camera_position(0.2, 2 * math.pi * k / 8, 0.3)) for k in range(8)]

    def sweep(n):
        for i in range(n):
            sim.shade(views[i % 8])

    ms = ctx.timed(sweep, 40, 8) / 40
    rec("config 5: LUT shading of 8192x16384 per view (64-wavelength LUT, 8 yaw angles)", 8192 * 16384, 8192 * 16384, 8, ms,
        f"LUT build 256x64 texels x 64 wavelengths incl. D2H: {t_lut * 1e3:.3f} ms per build (host wall clock)")
    if hasattr(sim, "shade_spectral"):
        ms = ctx.timed(lambda n: [sim.shade_spectral(views[i % 8], wl64) for i in range(n)], 8, 2) / 8
        rec("config 5 direct: per-cell 64-wavelength spectral evaluation of 8192x16384 (no LUT)", 8192 * 16384,
            8192 * 16384, 8, ms, "SFU/FP32-bound by construction (64 sin^2 per cell against 8 B)")
    sim.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="sbx", choices=["sbx", "reference"])
    ap.add_argument("--workload", default="sphere", choices=["sphere", "foam"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--bubbles", type=int, default=FOAM_BUBBLES)
    ap.add_argument("--rgb", default="lut", choices=["lut", "exact3", "none"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="theta rows per GPU (weak) or in total (strong)")
    ap.add_argument("--phi", type=int, default=NUM_PHI)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--other-configs", dest="other_configs", action="store_true", default=True)
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo rows: in-kernel peer-memory read (default) or NCCL send/recv per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_sbx(args)


if __name__ == "__main__":
    main()
