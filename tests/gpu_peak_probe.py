"""Runs the two integer-pipe peak kernels (k_imad_peak_cols / k_imad_peak_rows) a few times; meant to be wrapped in ncu:
    ncu --metrics <list in profiles/r02_imad_peak.md> --clock-control none -k regex:k_imad_peak python tests/gpu_peak_probe.py
and prints the event-timed rates next to it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zk-nullifier-sig_b200"))
import plume_b200

with plume_b200.PlumeContext(0) as ctx:
    plain, carry = ctx.measure_imad_rates(4096)
    print("plain IMAD.WIDE.U32 columns: %.4e limb products/s; carry-chain rows: %.4e limb products/s" % (plain, carry))
