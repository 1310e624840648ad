"""bench.py's bookkeeping that needs no GPU: the ncu traffic figure is tied to the kernel sources it was captured from, the
oracle is only reachable from the checker / CPU-baseline legs, and the reference arm prints the CUDA arm's config."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_py(code: str) -> str:
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stderr + r.stdout  # bench.py points fd 1 at stderr on import


def test_traffic_figure_is_refused_when_the_kernel_sources_changed(tmp_path):
    out = run_py("import bench, json; print('RESULT', json.dumps([bench.kernel_source_hash(), *bench.ncu_traffic()]))")
    sha, traffic, note = json.loads(re.search(r"RESULT (.*)", out).group(1))
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    if t.get("kernel_source_sha16") == sha:
        assert traffic == t["k_march_dram_bytes_per_launch"] and "stale" not in note
    else:  # the capture is of another kernel: no number, and the note says which hashes disagree
        assert traffic is None and "stale" in note and sha in note
    # a capture of other sources is refused
    out = run_py("import bench; bench.KERNEL_SOURCES = ('sbx_api.cu',); print('RESULT', bench.ncu_traffic()[0])")
    assert "RESULT None" in out


def test_bench_reaches_the_oracle_only_from_checker_and_baseline_legs():
    with open(os.path.join(ROOT, "bench.py")) as f:
        src = f.read()
    assert "sys.path.insert(0, os.path.join(ROOT, \"tests\"))" not in src  # helpers live in the package / oracle/ now
    uses = [m.start() for m in re.finditer(r"oracle_module\(\)", src)]
    owners = set()
    for pos in uses[1:]:  # the first hit is the definition
        head = src[:pos]
        owners.add(re.findall(r"\ndef (\w+)|\nclass (\w+)", head)[-1])
    names = {a or b for a, b in owners}
    assert names <= {"CpuReference", "check_sphere", "run_sbx"}, names  # run_sbx: the foam parity_check block only
    assert "oracle" not in open(os.path.join(ROOT, "soap_bubble_simulation_b200", "__init__.py")).read()


def test_reference_arm_prints_the_cuda_arms_config():
    code = ("import bench, argparse, json\n"
            "a = argparse.Namespace(workload='sphere', scaling='weak', rows=8192, phi=16384, rgb='lut', halo='peer', bubbles=4096)\n"
            "print('RESULT', json.dumps(bench.workload(a, 1)['config']))")
    cfg = json.loads(re.search(r"RESULT (.*)", run_py(code)).group(1))
    assert "8192x16384" in cfg["workload"] and cfg["grid_per_gpu"] == [8192, 16384]
    with open(os.path.join(ROOT, "bench.py")) as f:
        src = f.read()
    assert src.count('"config": wl["config"]') == 2  # both arms emit the same object


def test_parity_check_logic_accepts_the_oracle_and_rejects_a_wrong_row():
    """bench.py's check_sphere (the driver-visible parity gate of sharded runs) on CPU: fed with the whole-sphere oracle's
    own rows it must pass; with one band's boundary row shifted by one row (a stale ghost row) it must fail."""
    code = r'''
import bench, json, numpy as np
import soap_bubble_simulation_b200  # noqa: F401  (package import only: no device needed for the host-side helpers)
from soap_bubble_simulation_b200.synth import synthetic_field, dt_max
from soap_bubble_simulation_b200.sharding import band_partition
from soap_bubble_simulation_b200.drainage import View
O = bench.oracle_module()
nt, nphi, steps, world = 96, 64, 5, 3
dt = 0.25 * dt_max(nt, nphi)
view = View()
ora = O.OracleDrainage(nt, nphi)
ora.data[:] = synthetic_field(nt, nphi, dtype=np.float32).astype(np.float64)
ora.step(dt, steps)
full = ora.data.astype(np.float32)
px = O.shade_field(ora.data, nt, 0, 0.025, view.cam_pos, intensity=view.intensity, mode=1)
bands = band_partition(nt, world)
edges = [bench.sphere_edges(full[a:b], px[a:b]) for a, b in bands]
good = bench.check_sphere(view, nt, nphi, dt, steps, bands, edges, "unit test")
bad_field = full.copy()
a1 = bands[1][0]
bad_field[a1] = full[a1 + 1]  # rank 1's first row computed from a stale ghost row would look like this
edges_bad = [bench.sphere_edges(bad_field[a:b], px[a:b]) for a, b in bands]
bad = bench.check_sphere(view, nt, nphi, dt, steps, bands, edges_bad, "unit test")
print("RESULT", json.dumps([good["ok"], good["max_rel"], good["rgba_max_lsb"], good["rows_checked"], bad["ok"], bad["max_rel"]]))
'''
    out = run_py(code)
    ok, rel, lsb, rows, bad_ok, bad_rel = json.loads(re.search(r"RESULT (.*)", out).group(1))
    assert ok and rel <= 1.2e-7 and lsb <= 1 and rows == 3 * 2 * 8
    assert not bad_ok and bad_rel > 1e-5


def test_reference_arm_under_torchrun_runs_on_rank_zero_only():
    """`bench.py --impl reference` launched like the CUDA arm at N = 2: rank 0 alone times the CPU path and prints the one
    JSON line, the other rank exits 0 without work (no GPU, no rendezvous needed)."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_update_traffic_reads_the_ncu_summary(tmp_path):
    """scripts/update_traffic.py: the two dram__bytes sums of an ncu summary (with their units) become bytes per launch."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    try:
        code = ("import sys; sys.path.insert(0, 'scripts'); import update_traffic as u\n"
                "t = '  dram__bytes_read.sum      537.852672 Mbyte\\n  dram__bytes_write.sum   1.018144 Gbyte\\n'\n"
                "print('RESULT', int(round(u.metric(t, 'dram__bytes_read.sum') + u.metric(t, 'dram__bytes_write.sum'))))")
        out = run_py(code)
    finally:
        sys.path.pop(0)
    assert "RESULT 1555996672" in out
