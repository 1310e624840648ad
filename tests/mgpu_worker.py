"""torchrun worker for the multi-GPU parity test: latitude-band sharded drainage (+ fused colour) against
the single-domain CPU oracle.  Usage: torchrun --nproc-per-node N tests/mgpu_worker.py <nt> <nphi> <steps> <prec> <halo>
halo: nccl | peer (device-resident steps), host (sbx_drainage_step_host per rank, then device-resident steps),
foam (independent bubbles sharded by bubble index, no exchange)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib as O  # noqa: E402
import soap_bubble_simulation_b200 as S  # noqa: E402
from soap_bubble_simulation_b200.sharding import Comm, band_partition, bubble_range  # noqa: E402
from synth import dt_max, synthetic_field  # noqa: E402


def foam(nt, nphi, steps, prec, rank, world, local):
    """config 4 in small: 3 * world + 1 bubbles with their own h0 / radius, split by bubble_range, no exchange."""
    total = 3 * world + 1
    b0, b1 = bubble_range(total, world, rank)
    idx = np.arange(total) / (total - 1)
    h0, rad = (300.0 + 400.0 * idx) * 1e-9, 0.0125 + 0.0875 * idx
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, device=local, num_theta=nt, num_phi=nphi,
                              h_ref=1.0, batch=b1 - b0, batch_h0_m=h0[b0:b1], batch_radius_m=rad[b0:b1])
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    dt = 1e-3
    sim.step(dt, steps - 1)
    sim.step_fused(dt, view, 1)
    got, px = sim.download(), sim.download_rgb()
    ok, worst, worst_inc, worst_px = True, 0.0, 0.0, 0
    for k, b in enumerate(range(b0, b1)):
        ora = O.OracleDrainage(nt, nphi, h0_m=float(h0[b]), radius_m=float(rad[b]))
        ora.step(dt, steps)
        worst = max(worst, float(np.abs(got[k] / ora.data - 1).max()))
        dref = ora.data - h0[b]
        worst_inc = max(worst_inc, float(np.abs((got[k] - h0[b]) - dref).max() / np.abs(dref).max()))
        ref = O.shade_field(ora.data, nt, 0, float(rad[b]), view.cam_pos, intensity=4.0, mode=1)
        worst_px = max(worst_px, int(np.abs(px[k].astype(int) - ref.astype(int)).max()))
    ok = worst <= 1e-5 and worst_inc <= 1e-3 and worst_px <= 1
    allok = [None] * world
    dist.all_gather_object(allok, (b0, b1, ok, worst, worst_inc, worst_px))
    covered = sorted(b for r in allok for b in range(r[0], r[1]))
    ok_all = all(r[2] for r in allok) and covered == list(range(total))
    if rank == 0:
        print(f"MGPU world={world} foam {total} bubbles of {nt}x{nphi}: err={max(r[3] for r in allok):.3e} "
              f"incr={max(r[4] for r in allok):.3e} rgba_max_diff={max(r[5] for r in allok)} ok={ok_all}", flush=True)
    return ok_all


def wgsl(nt, nphi, steps, halo, rank, world, local):
    """WGSL semantics (thickness + concentration, Marangoni on) sharded by latitude bands: both fields' halo rows by
    NCCL send/recv ("wgsl_nccl") or pushed from inside k_wgsl_tma over peer memory ("wgsl_peer")."""
    rb, re = band_partition(nt, world)[rank]
    sim = S.GPUDrainageSimulator(500e-9, nphi, nt, device=local, row_begin=rb, row_end=re)
    rng = np.random.default_rng(3)
    hf = synthetic_field(nt, nphi, h0_nm=5000.0, noise_nm=20.0).astype(np.float32)
    cf = np.clip(0.5 + 0.2 * rng.standard_normal((nt, nphi)), 0, 1).astype(np.float32)
    sim.set_marangoni_enabled(True)
    sim.upload(hf[rb:re])
    sim.upload_concentration(cf[rb:re])
    comm = Comm(rank, world, local)
    sim.attach_comm(comm)
    if halo == "wgsl_peer":
        comm.enable_peer_halo(sim)
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    sim.dispatch(steps - 1)
    sim.dispatch(1, view)
    gh, gc, px = sim.download(np.float32)[0], sim.download_concentration()[0], sim.download_rgb()[0]
    bad = Comm.peer_status(sim) if halo == "wgsl_peer" else 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (gh, gc, px, bad))
    ok = True
    if rank == 0:
        p = O.wgsl_params(grid_width=nphi, grid_height=nt, marangoni_enabled=1, dt=0.001)
        ho, co = hf, cf
        for _ in range(steps):
            ho, co = O.wgsl_step(p, ho, co)
        fh = np.concatenate([g[0] for g in gathered])
        fc = np.concatenate([g[1] for g in gathered])
        fpx = np.concatenate([g[2] for g in gathered])
        eh = float(np.abs(fh.astype(np.float64) / ho - 1).max())
        ec = float(np.abs(fc.astype(np.float64) - co).max())
        ref = O.shade_field(ho.astype(np.float64), nt, 0, 0.025, view.cam_pos, intensity=4.0, mode=1)
        dpx = int(np.abs(fpx.astype(int) - ref.astype(int)).max())
        ok = eh <= 1e-5 and ec <= 1e-5 and dpx <= 1 and not any(g[3] for g in gathered)
        print(f"MGPU world={world} wgsl grid={nt}x{nphi} halo={halo}: err_h={eh:.3e} err_c={ec:.3e} rgba_max_diff={dpx} ok={ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    comm.close()
    return int(flag.item()) == 1


def main():
    nt, nphi, steps, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    halo = sys.argv[5] if len(sys.argv) > 5 else "nccl"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if halo == "foam" or halo.startswith("wgsl"):
        ok = foam(nt, nphi, steps, prec, rank, world, local) if halo == "foam" else wgsl(nt, nphi, steps, halo, rank, world, local)
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    rb, re = band_partition(nt, world)[rank]
    sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=prec, device=local, num_theta=nt,
                              num_phi=nphi, h_ref=0.0, row_begin=rb, row_end=re)
    band = synthetic_field(nt, nphi, row0=rb, rows=re - rb)
    if prec == S.PREC_F32:
        band = band.astype(np.float32).astype(np.float64)
    sim.upload(band)
    comm = Comm(rank, world, local)
    sim.attach_comm(comm)
    if halo in ("peer", "host", "peer_mix", "peer_reset"):
        comm.enable_peer_halo(sim)
    dt = min(5e-3, 0.25 * dt_max(nt, nphi))  # 5 ms steps reach the first TFE front (t = 0.5 s) quickly
    view = S.View(mode=S.RGB_LUT, format=S.OUT_RGBA8)
    if halo == "host":
        # the end-to-end path: every rank hands its band in host memory to sbx_drainage_step_host (chunk-pipelined,
        # input halo rows exchanged inside the call) and feeds the result back; then device-resident steps follow,
        # whose ghost rows have to be fetched first
        cur = band.astype(np.float32)[None].copy()
        out_h = np.empty_like(cur)
        out_px = np.empty(cur.shape + (4,), np.uint8)
        n_host = steps - 4
        for _ in range(n_host):
            sim.step_host(dt, 1, view, cur, out_h, out_px)
            cur, out_h = out_h, cur
        assert np.array_equal(sim.download(np.float32), cur)
        sim.step(dt, 3)
        sim.step_fused(dt, view, 1)
    elif halo == "peer_reset":
        # reset on every rank between peer-halo steps: the ranks leave reset() at different times, and a neighbour that
        # is ahead must not read this band's pre-reset rows (nor this band overwrite rows a neighbour is still reading)
        sim.step(dt, 7)
        sim.reset(5e-7)
        sim.step(dt, steps - 1)
        sim.step_fused(dt, view, 1)
    elif halo == "peer_mix":
        # peer-halo steps read the neighbours' boundary rows in place; everything else that rewrites the field between
        # them (upload, the host-buffer step) has to meet the neighbours first and hand fresh ghost rows to whoever
        # wants them afterwards
        a = steps // 3
        sim.step(dt, a)
        cur = sim.download(np.float32)
        sim.upload(cur[0])                      # same values: the trajectory is unchanged
        sim.step(dt, a)
        cur = sim.download(np.float32).copy()
        out_h = np.empty_like(cur)
        for _ in range(2):
            sim.step_host(dt, 1, view, cur, out_h, None)
            cur, out_h = out_h, cur
        sim.step(dt, steps - 2 * a - 3)
        sim.step_fused(dt, view, 1)
    else:
        sim.step(dt, steps - 1)
        sim.step_fused(dt, view, 1)
    got = sim.download()[0]
    px = sim.download_rgb()[0]
    mn, mx = sim.minmax_global()
    peer_bad = Comm.peer_status(sim) if halo in ("peer", "host", "peer_mix", "peer_reset") else 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (rb, re, got, px, sim.fronts(), peer_bad))
    ok = True
    if rank == 0:
        full = np.concatenate([g[2] for g in gathered])
        frame = np.concatenate([g[3] for g in gathered])
        ora = O.OracleDrainage(nt, nphi)
        f0 = synthetic_field(nt, nphi)
        if prec == S.PREC_F32:
            f0 = f0.astype(np.float32).astype(np.float64)
        ora.data[:] = 5e-7 if halo == "peer_reset" else f0  # reset(): uniform film, clock and fronts start over
        ora.step(dt, steps)
        if prec == S.PREC_F64_STRICT:
            ok = np.array_equal(full, ora.data)
            err = float(np.abs(full - ora.data).max())
        else:
            err = float(np.abs(full / ora.data - 1).max())
            ok = err <= (1e-5 if prec == S.PREC_F32 else 1e-12)  # f32: ~6e-8 per step accumulates
        ok = ok and gathered[0][4] == ora.fronts() and not any(g[5] for g in gathered)
        ok = ok and abs(mn - ora.min_thickness()) <= 1e-5 * mn and abs(mx - ora.max_thickness()) <= 1e-5 * mx
        ref = O.shade_field(ora.data, nt, 0, 0.025, view.cam_pos, intensity=4.0, mode=1)
        dpx = int(np.abs(frame.astype(int) - ref.astype(int)).max())
        ok = ok and dpx <= 1
        print(f"MGPU world={world} grid={nt}x{nphi} prec={prec} halo={halo}: err={err:.3e} rgba_max_diff={dpx} ok={ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
