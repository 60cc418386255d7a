#!/usr/bin/env python
"""Occupancy statistics of the synthetic scenes: N returns, F occupied voxel_size cells, K occupied sample_voxel_size
cells per sweep (what SubSampleFrame / grid sampling keep), range percentiles, ground share.
usage: python tools/scene_stats.py [street|suburb] [HDL64|HDL64E|...] [frames]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_icp_b200 import synthetic as syn   # noqa: E402

profile = sys.argv[1] if len(sys.argv) > 1 else "suburb"
sensor = getattr(syn, sys.argv[2] if len(sys.argv) > 2 else "HDL64E")
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
scene = syn.UrbanScene(1234, profile=profile)
print("scene %s: %d boxes, %d cylinders; sensor %s" % (profile, len(scene.boxes), len(scene.cyl), sensor.name))
for s in syn.make_sequence(frames, sensor, seed=1234, start=30, scene=scene):
    x = s["xyz"]
    r = np.linalg.norm(x, axis=1)
    F = len(np.unique(np.trunc(x / 0.5).astype(np.int64), axis=0))
    K = len(np.unique(np.trunc(x / 1.5).astype(np.int64), axis=0))
    ground = float((np.abs(x[:, 2] + sensor.height) < 0.15).mean())
    print("frame %d: N %d  F(0.5 m) %d  K(1.5 m) %d  range p10/p50/p90 %s  ground share %.2f"
          % (s["frame_idx"], len(x), F, K, np.percentile(r, [10, 50, 90]).round(1), ground))
