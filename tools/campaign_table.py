#!/usr/bin/env python3
"""Reproduce the SHAPE of the reference's published fault-injection table (docs/source/results/msp430.rst,
5 000 injections per cell, "OK / Err / DWC-detected") with the on-device injector.  Context, not a parity claim:
the MSP430 campaigns flip random registers/RAM of a whole program; here every flip hits a LIVE replica value."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coast_b200 as cb  # noqa: E402
from coast_b200 import campaign as cp  # noqa: E402

rt = cb.Runtime(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rows = []
t0 = time.time()
for mode, name in (("", "Unmitigated"), ("-DWC", "-DWC"), ("-TMR", "-TMR"), ("-TMR -countErrors", "-TMR -countErrors")):
    cells = []
    for wl in ("mm", "crc16", "qsort", "sha256", "aes"):
        s, _ = cp.run_campaign(rt, wl, mode, n, seed=7)
        cells.append(s.row())
    rows.append((name, cells))
print(f"| Config | MxM (exact int) | CRC16 | QS (580 ints) | SHA-256 | AES-128 |  ({n} single-bit flips of live replica values per cell, {time.time() - t0:.1f} s total)")
print("|---|---|---|---|---|---|")
for name, cells in rows:
    print(f"| {name} | " + " | ".join(cells) + " |")
