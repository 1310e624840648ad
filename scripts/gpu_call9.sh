#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/pcie_probe_hugepage.py 2 > $OUT/r02_pcie_probe_hugepage_n2.txt 2>&1; cat $OUT/r02_pcie_probe_hugepage_n2.txt
grep -i huge /proc/meminfo; cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 300 python scripts/resident_frame_probe.py > $OUT/r02_resident_frame_probe.txt 2>&1; cat $OUT/r02_resident_frame_probe.txt
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "spectral" 2>&1 | tail -3
timeout 300 python - <<'PY'
import sys, math, numpy as np, torch
sys.path.insert(0, '.')
import soap_bubble_simulation_b200 as S
from soap_bubble_simulation_b200.synth import synthetic_field
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sim = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=8192, num_phi=16384, h_ref=0.0)
sim.upload(synthetic_field(8192, 16384, dtype=np.float32)); sim.set_stream(stream.cuda_stream)
views = [S.View(cam_pos=S.camera_position(0.2, 2 * math.pi * k / 8, 0.3)) for k in range(8)]
for nl in (7, 64):
    wl = 380.0 + 400.0 * np.arange(nl) / max(nl - 1, 1)
    sim.shade_spectral(views[0], wl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(8): sim.shade_spectral(views[i], None)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    print(f"spectral direct {nl} wavelengths: {ms:.4f} ms per view, {8192*16384/ms/1e6:.1f} Gcell/s, {8192*16384*nl/ms/1e9:.2f} T wavelength-evaluations/s")
PY
