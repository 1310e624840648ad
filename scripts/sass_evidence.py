#!/usr/bin/env python
"""SASS excerpt of the headline kernel for profiles/ (no GPU needed: cuobjdump reads the built libsbx.so).

    python scripts/sass_evidence.py > profiles/r02_sass_k_march_tma.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "soap_bubble_simulation_b200", "csrc", "libsbx.so")
FUN = "_ZN3sbx11k_march_tmaIfLi1ELb0ELi1EEEvNS_8StepArgsIT_EE"  # k_march_tma<float, SBX_RGB_LUT, false>

txt = subprocess.run(["cuobjdump", "-sass", "-fun", FUN, SO], capture_output=True, text=True).stdout
ins = [re.sub(r"\s+/\*.*", "", re.sub(r"^\s+/\*([0-9a-f]{4})\*/\s+", r"\1  ", l))
       for l in txt.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
ops = collections.Counter()
for l in ins:
    m = re.match(r"[0-9a-f]{4}\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(1).split(".")[0]] += 1
res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout.splitlines()
k = next(i for i, l in enumerate(res) if FUN in l)
print("SASS evidence for the headline kernel k_march_tma<float, SBX_RGB_LUT, TFE=false> (sm_100a), from")
print(f"  cuobjdump -sass -fun {FUN} soap_bubble_simulation_b200/csrc/libsbx.so")
print("  cuobjdump -res-usage soap_bubble_simulation_b200/csrc/libsbx.so")
print("(regenerate with scripts/sass_evidence.py after `make -C soap_bubble_simulation_b200/csrc`)\n")
print("resource usage:\n  " + res[k].strip() + "\n  " + res[k + 1].strip() + "\n")
print(f"instructions: {len(ins)}")
print("opcode histogram (top 30): " + ", ".join(f"{a} {b}" for a, b in ops.most_common(30)) + "\n")
print("""What to look for:
  UBLKCP.S.G           cp.async.bulk global -> shared: the TMA bulk copies of the strip rows (the form with desc[..] carries
                       the L2 evict_last policy of the band-edge rows)
  SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT    mbarrier expect_tx / arrive / try_wait of the ring
  STS [R..], R..  immediately before  SYNCS.ARRIVE.TRANS64.A1T0    the stage hand-back: store of the OR of the loaded
                       registers to the per-warp sink word, then the arrive (sbx_tma.cuh: mbar_arrive_after)
  TEX.LL               the LUT fetch through the texture unit (4 per row and thread)
  LDS.128 / STG.E.128  16-byte shared loads of the row packets; 16-byte stores of h' and of 4 RGBA8 pixels
  the LDG.E are the per-band coefficient / view tables (prologue) and the pole partial sums (epilogue): the thickness
  field itself reaches the SM only through UBLKCP
""")
print("memory / synchronisation instructions in program order:")
for l in ins:
    if re.search(r"UBLKCP|SYNCS|TEX|STG\.E|LDS\.128|STS |LDG|MEMBAR|ATOM|RED\.|LD\.E|ST\.E|ERRBAR|DEPBAR", l):
        print("  " + l)
