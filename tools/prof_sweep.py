#!/usr/bin/env python
"""BASELINE.json configs[4] (replicate sweep) on one GPU: bench.py's replicate_sweep block on its own, for launch lists
and quick A/B runs.   python tools/prof_sweep.py [R]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import pedigreesim_b200 as ps  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
torch.cuda.set_device(0)
print(json.dumps(bench.replicate_sweep(ps, torch, 0, 0, R)))
