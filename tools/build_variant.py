"""Builds a tuning variant of libsnacb200.so: python tools/build_variant.py NAME -DSNAC_RING_RD=12 ...
-> snac_b200/variants/libsnacb200_NAME.so; run with SNAC_B200_LIB=<that path>."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from snac_b200.build import HERE, build_library

name, defs = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
print(build_library(defines=defs, out=os.path.join(HERE, "variants", f"libsnacb200_{name}.so")))
