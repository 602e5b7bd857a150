"""Build a variant of libocn_b200.so with extra nvcc defines, for A/B runs inside one gpurun call:

    python scripts/build_variant.py minb6 -DOCN_SEG_MINB=6      ->  ocn_b200/variants/libocn_b200_minb6.so
    OCN_B200_LIB=ocn_b200/variants/libocn_b200_minb6.so python scripts/ab_hub.py
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocn_b200 import build as b  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(b.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libocn_b200_{name}.so")
cmd = ["nvcc"] + b.NVCC_FLAGS + flags + ["-o", out] + b.sources()
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    sys.exit(res.stdout + res.stderr)
print(out)
