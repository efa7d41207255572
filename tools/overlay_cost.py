"""What the 16 blended overlay quads of configs[2] cost: mesh4k_tex with and without them (and, without them,
deferred against the single kernel: DTR_B200_DEFER=0).  usage: python tools/overlay_cost.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ov = int(os.environ.get("OVERLAYS", "16"))
bench.WORKLOADS["mesh4k_tex"]["overlays"] = ov
env = bench.Env(0, 1, 0)
out = bench.measure_views(env, "mesh4k_tex", 32, 20, 5, 0, False)
r = out["roofline"]
print(f"overlays {ov} DEFER={os.environ.get('DTR_B200_DEFER', '1')}: ms/step {out['ms_per_step']:.4f} raster {r['ms_per_launch']:.4f} "
      f"kernels {json.dumps({k: round(v, 4) for k, v in r['raster_kernels_ms_per_step'].items()})} parity {out['parity_checked']}")
