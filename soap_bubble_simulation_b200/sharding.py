"""Latitude-band sharding across the GPUs of one box: one process per GPU, torch.distributed for the
rendezvous, NCCL send/recv (inside libsbx) for the per-step halo rows.

The sphere is cut into contiguous bands of theta rows; phi stays whole on every rank, so the periodic
seam never crosses devices.  The pole rows and the rings whose mean they take live on the first and last
rank.  TFE front state is a deterministic function of time and is replicated, not communicated.
Independent foam bubbles shard by bubble index with no exchange at all (``bubble_range``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _lib as L


def band_partition(num_theta: int, world_size: int) -> List[Tuple[int, int]]:
    """Row ranges [begin, end) per rank: as even as possible, every band >= 2 rows so that a pole row is
    never separated from the ring next to it."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if num_theta < 2 * world_size or num_theta < 3:
        raise ValueError(f"{num_theta} theta rows cannot be split into {world_size} bands of >= 2 rows")
    base, rem = divmod(num_theta, world_size)
    out, r0 = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((r0, r0 + n))
        r0 += n
    return out


def bubble_range(num_bubbles: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Independent bubbles: contiguous slice of the batch for this rank (no data-path collective)."""
    base, rem = divmod(num_bubbles, world_size)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def halo_plan(rank: int, world_size: int, rows_local: int):
    """The per-step exchange libsbx performs (sbx_comm_halo), in local-row terms: owned rows are 1..rows_local,
    ghost rows 0 and rows_local+1.  Returns [(send_local_row, peer, recv_ghost_row)]: the row sent to `peer` and
    the ghost row filled from the same peer."""
    plan = []
    if rank > 0:
        plan.append((1, rank - 1, 0))
    if rank < world_size - 1:
        plan.append((rows_local, rank + 1, rows_local + 1))
    return plan


def owns_poles(row_begin: int, row_end: int, num_theta: int):
    """(has_top, has_bottom): which pole rows -- and therefore which ring means -- this band computes."""
    return row_begin == 0, row_end == num_theta


class Comm:
    """sbx_comm: an NCCL communicator owned by libsbx.  The 128-byte unique id is created on rank 0 and
    handed to the other ranks through torch.distributed (any backend: gloo on CPU hosts, nccl on GPUs)."""

    def __init__(self, rank: int, world_size: int, device: int):
        import torch
        import torch.distributed as dist

        self.rank, self.world_size = rank, world_size
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            L.check(L.lib().sbx_comm_unique_id(ident.ctypes.data))
        if world_size > 1:
            box = [ident.tobytes()]
            dist.broadcast_object_list(box, src=0)
            ident = np.frombuffer(box[0], dtype=np.uint8).copy()
        self._c = C.c_void_p()
        L.check(L.lib().sbx_comm_init(ident.ctypes.data, world_size, rank, device, C.byref(self._c)))
        del torch

    def enable_peer_halo(self, sim) -> None:
        """Switch `sim` (already attached to this communicator) from per-step NCCL send/recv to the in-kernel
        peer-memory halo (the stencil kernel reads the neighbours' boundary rows in place over NVLink): exchange CUDA IPC descriptors of the field buffers with the two latitude
        neighbours and map them (sbx_drainage_ipc_export / _attach)."""
        import torch.distributed as dist

        blob = np.zeros(L.IPC_BYTES, dtype=np.uint8)
        L.check(L.lib().sbx_drainage_ipc_export(sim._h, blob.ctypes.data))
        blobs = [None] * self.world_size
        if self.world_size > 1:
            dist.all_gather_object(blobs, blob.tobytes())
        else:
            blobs[0] = blob.tobytes()
        up = np.frombuffer(blobs[self.rank - 1], dtype=np.uint8).copy() if self.rank > 0 else None
        dn = np.frombuffer(blobs[self.rank + 1], dtype=np.uint8).copy() if self.rank < self.world_size - 1 else None
        L.check(L.lib().sbx_drainage_ipc_attach(sim._h, up.ctypes.data if up is not None else None,
                                                dn.ctypes.data if dn is not None else None))
        if self.world_size > 1:
            dist.barrier()  # nobody steps before every neighbour has mapped every buffer

    @staticmethod
    def peer_status(sim) -> int:
        out = C.c_int32()
        L.check(L.lib().sbx_drainage_peer_status(sim._h, C.byref(out)))
        return out.value

    def close(self) -> None:
        if self._c:
            L.lib().sbx_comm_destroy(self._c)
            self._c = C.c_void_p()
