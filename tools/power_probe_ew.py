"""Bandwidth-bound BatchNorm kernels at several grid sizes (UNETK_EW_BLOCKS_PER_SM): GB/s, W, pJ/B."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for per_sm in (8, 6, 4, 3, 2):
    env = dict(os.environ, UNETK_EW_BLOCKS_PER_SM=str(per_sm), ONLY_EW="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "power_probe.py"), "--seconds", "2"], env=env, capture_output=True, text=True).stdout
    for line in out.splitlines():
        if line.startswith(("bnrelu_apply", "bn_bwd")):
            print(f"blocks/SM {per_sm}: {line}", flush=True)
