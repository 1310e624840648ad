"""Host-side logic of the latitude-band sharding, exercised on CPU with torch.distributed (gloo, world_size 2
and 3): partitioning, the halo plan libsbx executes over NCCL, pole ownership, and the NCCL-id rendezvous.
A numpy band stepper (test-only, same formula as the oracle, drainage.rs:314-381,480-503) stands in for the
kernels so that the exchange schedule itself can be checked against the single-domain oracle."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from synth import dt_max, synthetic_field


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def band_step_numpy(buf, row_begin, num_theta, num_phi, dt, p):
    """One CPU-semantics step of the owned rows of `buf` ((rows_local+2) x num_phi incl. ghosts), TFE off."""
    rows_local = buf.shape[0] - 2
    out = buf.copy()
    dth, dph = math.pi / (num_theta - 1), 2.0 * math.pi / num_phi
    coef = p["density"] * p["gravity"] / (3.0 * p["viscosity"])
    for l in range(1, rows_local + 1):
        g = row_begin + l - 1
        if g == 0 or g == num_theta - 1:
            continue
        th = g * dth
        s, c = math.sin(th), math.cos(th)
        h = buf[l]
        d2t = (buf[l + 1] - 2.0 * h + buf[l - 1]) / (dth * dth)
        dht = (buf[l + 1] - buf[l - 1]) / (2.0 * dth)
        d2p = (np.roll(h, -1) - 2.0 * h + np.roll(h, 1)) / (dph * dph)
        lap = (d2t + c / s * dht + d2p / (s * s)) / (p["radius_m"] ** 2)
        nh = np.maximum(h + dt * (-coef * (h * h * h) * s + p["diffusion"] * lap), 0.0)
        out[l] = np.where(h < p["critical_thickness_m"], h, nh)
    if row_begin == 0:
        out[1] = out[2].sum() / num_phi if False else np.add.reduce(out[2]) / num_phi
    if row_begin + rows_local == num_theta:
        out[rows_local] = np.add.reduce(out[rows_local - 1]) / num_phi
    return out


def _worker(rank, world, port, nt, nphi, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from soap_bubble_simulation_b200 import _lib as L
    from soap_bubble_simulation_b200.sharding import band_partition, halo_plan, owns_poles

    rb, re = band_partition(nt, world)[rank]
    rows = re - rb
    full = synthetic_field(nt, nphi)
    band = synthetic_field(nt, nphi, row0=rb, rows=rows)
    assert np.array_equal(band, full[rb:re])
    buf = np.zeros((rows + 2, nphi))
    buf[1:-1] = band
    dt = 0.25 * dt_max(nt, nphi)

    def exchange(b):
        t = torch.from_numpy(b)
        ops = []
        recv = {}
        for send_row, peer, ghost_row in halo_plan(rank, world, rows):
            ops.append(dist.P2POp(dist.isend, t[send_row].clone(), peer))
            recv[ghost_row] = torch.empty(nphi, dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, recv[ghost_row], peer))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for g, v in recv.items():
            t[g] = v

    exchange(buf)
    for _ in range(steps):
        buf = band_step_numpy(buf, rb, nt, nphi, dt, O.DEFAULTS)
        exchange(buf)
    # NCCL unique id rendezvous used by sharding.Comm: rank 0 creates, everyone receives the same 128 bytes
    ident = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        L.check(L.lib().sbx_comm_unique_id(ident.ctypes.data))
    box = [ident.tobytes()]
    dist.broadcast_object_list(box, src=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rb, re, owns_poles(rb, re, nt), box[0], buf[1:-1].copy()))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nt,nphi", [(2, 24, 32), (3, 31, 20)])
def test_band_sharding_with_halo_exchange_matches_single_domain(world, nt, nphi):
    steps = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nt, nphi, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ora = O.OracleDrainage(nt, nphi)
    ora.set_marginal_regeneration(False)
    ora.data[:] = synthetic_field(nt, nphi)
    ora.step(0.25 * dt_max(nt, nphi), steps)
    rows = np.concatenate([g[4] for g in gathered])
    assert rows.shape == (nt, nphi)
    assert np.abs(rows / ora.data - 1).max() < 1e-13
    assert gathered[0][2] == (True, False) and gathered[-1][2] == (False, True)
    assert all(g[3] == gathered[0][3] and len(g[3]) == 128 for g in gathered)
    assert [g[0] for g in gathered][0] == 0 and gathered[-1][1] == nt


def test_band_partition_properties():
    from soap_bubble_simulation_b200.sharding import band_partition, bubble_range, halo_plan

    for nt in (3, 8, 129, 8192, 65536):
        for w in (1, 2, 3, 4, 8):
            if nt < 2 * w or nt < 3:
                with pytest.raises(ValueError):
                    band_partition(nt, w)
                continue
            parts = band_partition(nt, w)
            assert parts[0][0] == 0 and parts[-1][1] == nt
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [e - b for b, e in parts]
            assert min(sizes) >= 2 and max(sizes) - min(sizes) <= 1
    assert band_partition(8192 * 8, 8)[3] == (3 * 8192, 4 * 8192)
    got = [bubble_range(4096, 8, r) for r in range(8)]
    assert got[0] == (0, 512) and got[-1] == (3584, 4096)
    assert sum(e - b for b, e in [bubble_range(10, 4, r) for r in range(4)]) == 10
    assert halo_plan(0, 1, 5) == []
    assert halo_plan(0, 2, 5) == [(5, 1, 6)]
    assert halo_plan(1, 3, 4) == [(1, 0, 0), (4, 2, 5)]
