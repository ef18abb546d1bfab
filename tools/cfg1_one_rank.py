#!/usr/bin/env python
"""configs[0] as ONE rank: a few steps through the ABI (for an ncu launch list: how much of the 0.089 ms step is kernel time)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from specfem3d_b200 import abi
from specfem3d_b200.configs import cfg1

print(bench.time_config(abi.load(), cfg1(1, 1, 80), ("elastic_forces", "newmark"), steps=20, warmup=5))
