#!/usr/bin/env python
"""Does the host side of the copies (IOMMU / page size) set the aggregate D2H ceiling?  Same copies as pcie_probe.py with
three kinds of pinned host memory: cudaHostAlloc (torch pin_memory), mmap + MADV_HUGEPAGE + cudaHostRegister, and
mmap(MAP_HUGETLB) + cudaHostRegister when the box has reserved huge pages."""
import ctypes
import json
import mmap
import sys
import time

import torch

MB = 1 << 20
SIZE = 256 * MB
REPS = 6
libc = ctypes.CDLL("libc.so.6", use_errno=True)
libc.mmap.restype = ctypes.c_void_p
libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
MADV_HUGEPAGE = 14
MAP_HUGETLB = 0x40000


def host_buffer(kind):
    if kind == "cudaHostAlloc":
        return torch.empty(SIZE, dtype=torch.uint8, pin_memory=True)
    flags = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | (MAP_HUGETLB if kind == "hugetlb" else 0)
    p = libc.mmap(None, SIZE, mmap.PROT_READ | mmap.PROT_WRITE, flags, -1, 0)
    if p in (None, ctypes.c_void_p(-1).value):
        raise OSError(ctypes.get_errno(), "mmap failed")
    if kind == "thp":
        libc.madvise(ctypes.c_void_p(p), SIZE, MADV_HUGEPAGE)
    ctypes.memset(ctypes.c_void_p(p), 1, SIZE)  # fault the pages in
    rc = torch.cuda.cudart().cudaHostRegister(p, SIZE, 1)  # cudaHostRegisterPortable
    if int(rc) != 0:
        raise RuntimeError(f"cudaHostRegister -> {rc}")
    arr = (ctypes.c_uint8 * SIZE).from_address(p)
    return torch.frombuffer(arr, dtype=torch.uint8)


def main():
    n = min(torch.cuda.device_count(), int(sys.argv[1]) if len(sys.argv) > 1 else 8)
    for d in range(n):
        torch.cuda.set_device(d)
        torch.cuda.synchronize(d)
    for kind in ("cudaHostAlloc", "thp", "hugetlb"):
        try:
            hin = [host_buffer(kind) for _ in range(n)]
            hout = [host_buffer(kind) for _ in range(n)]
        except Exception as e:
            print(json.dumps({"host_memory": kind, "unavailable": str(e)}), flush=True)
            continue
        da = [torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
        db = [torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
        si = [torch.cuda.Stream(device=d) for d in range(n)]
        so = [torch.cuda.Stream(device=d) for d in range(n)]

        def run(devs, h2d, d2h):
            for rep in range(2):
                for d in range(n):
                    torch.cuda.synchronize(d)
                t0 = time.perf_counter()
                for _ in range(REPS):
                    for d in devs:
                        if h2d:
                            with torch.cuda.stream(si[d]):
                                da[d].copy_(hin[d], non_blocking=True)
                        if d2h:
                            with torch.cuda.stream(so[d]):
                                hout[d].copy_(db[d], non_blocking=True)
                for d in range(n):
                    torch.cuda.synchronize(d)
                el = time.perf_counter() - t0
            return round(SIZE * REPS * len(devs) / el / 1e9, 2)

        devs = list(range(n))
        print(json.dumps({"host_memory": kind, "gpus": devs, "one_gpu_d2h_GBps": run([0], False, True),
                          "aggregate_h2d_GBps": run(devs, True, False), "aggregate_d2h_GBps": run(devs, False, True),
                          "aggregate_bidir_each_direction_GBps": run(devs, True, True)}), flush=True)
        del hin, hout, da, db


if __name__ == "__main__":
    main()
