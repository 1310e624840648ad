#!/usr/bin/env python
"""Host <-> device copy ceilings of the box, for reading the end-to-end (host-buffer) numbers against.

One process, every visible GPU: pinned host buffers (cudaHostAlloc through torch), 256 MiB per copy,
H2D alone, D2H alone, both directions at once -- first one GPU at a time, then all GPUs concurrently.
Prints one JSON line per measurement; results are committed under profiles/.
"""
import json
import sys
import time

import torch

MB = 1 << 20
SIZE = 256 * MB
REPS = 6


def main():
    n = torch.cuda.device_count()
    only = int(sys.argv[1]) if len(sys.argv) > 1 else n
    n = min(n, only)
    host_in = [torch.empty(SIZE, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    host_out = [torch.empty(SIZE, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
    dev_a, dev_b, s_in, s_out = [], [], [], []
    for d in range(n):
        with torch.cuda.device(d):
            dev_a.append(torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{d}"))
            dev_b.append(torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{d}"))
            s_in.append(torch.cuda.Stream(device=d))
            s_out.append(torch.cuda.Stream(device=d))

    def sync_all():
        for d in range(n):
            torch.cuda.synchronize(d)

    def run(devs, h2d, d2h):
        for rep in range(2):  # first pass warms up
            sync_all()
            t0 = time.perf_counter()
            for _ in range(REPS):
                for d in devs:
                    if h2d:
                        with torch.cuda.stream(s_in[d]):
                            dev_a[d].copy_(host_in[d], non_blocking=True)
                    if d2h:
                        with torch.cuda.stream(s_out[d]):
                            host_out[d].copy_(dev_b[d], non_blocking=True)
            sync_all()
            el = time.perf_counter() - t0
        per_dir = SIZE * REPS * len(devs) / el / 1e9
        return per_dir

    for d in range(n):
        r = {"gpus": [d], "h2d_GBps": round(run([d], True, False), 2), "d2h_GBps": round(run([d], False, True), 2),
             "bidir_each_direction_GBps": round(run([d], True, True), 2)}
        print(json.dumps(r), flush=True)
    for k in (2, 4, 8):
        if k > n:
            break
        devs = list(range(k))
        r = {"gpus": devs, "aggregate_h2d_GBps": round(run(devs, True, False), 2),
             "aggregate_d2h_GBps": round(run(devs, False, True), 2),
             "aggregate_bidir_each_direction_GBps": round(run(devs, True, True), 2)}
        print(json.dumps(r), flush=True)
    if n >= 2:  # neighbours on a shared switch?  pairs (0,1) vs (0, n-1)
        for pair in ([0, 1], [0, n - 1], [0, n // 2]):
            r = {"gpus": pair, "aggregate_d2h_GBps": round(run(pair, False, True), 2)}
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
