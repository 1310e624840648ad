"""Runs on the GPU box right after scripts/profile.sh: put the dram bytes of the fresh ncu capture and the hash of the kernel
sources it was taken from into profiles/traffic.json (bench.py refuses a capture of other sources), and leave a copy under
gpurun_out/ so that it travels back.

    python scripts/update_traffic.py gpurun_out/r02_k_march_tma_lut.txt
"""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_hash)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def metric(text, name):
    m = re.search(rf"{re.escape(name)}\s+([0-9.]+)\s+(\w+)", text)
    if not m:
        raise SystemExit(f"{name} not found in the ncu summary")
    return float(m.group(1)) * UNIT[m.group(2)]


def main():
    summary = sys.argv[1]
    text = open(summary).read()
    rd, wr = metric(text, "dram__bytes_read.sum"), metric(text, "dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(path))
    t["k_march_dram_bytes_per_launch"] = int(round(rd + wr))
    t["kernel_source_sha16"] = bench.kernel_source_hash()
    t["capture"] = re.sub(r"dram__bytes_read\.sum [0-9.]+ MB \+ dram__bytes_write\.sum [0-9.]+ MB",
                          f"dram__bytes_read.sum {rd / 1e6:.3f} MB + dram__bytes_write.sum {wr / 1e6:.3f} MB", t["capture"])
    with open(path, "w") as f:
        json.dump(t, f, indent=1)
        f.write("\n")
    shutil.copy(path, os.path.join(ROOT, "gpurun_out", "traffic.json"))
    print(json.dumps(t))


if __name__ == "__main__":
    main()
