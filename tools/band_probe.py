"""One rank's share of the configs[3] band split, on ONE GPU: the context rasterises only the band an N-rank split would
give rank 0 (geometry replicated), so the raster kernel of a small launch can be timed without an N-GPU box.
usage: DTR_B200_LIB=... python tools/band_probe.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dtrenderer_b200 import api, multigpu, scenes  # noqa: E402

w, h, n = 3840, 2160, 1_000_000
p, color = scenes.small_triangles(w, h, n, seed=7)
for world in [int(a) for a in sys.argv[1:]] or [8]:
    r = api.Renderer(w, h, 1, 0)
    y0, y1 = multigpu.band_rows(h, world, 0, r.tile_height())
    if world > 1:
        r.set_band(y0, y1)
    r.begin_frame(0)
    r.clear((0, 0, 0))
    r.triangles(p, color, scenes.DEFAULT_TRIANGLE_TRANSFORM)
    r.flush()
    for _ in range(5):
        r.replay()
    r.set_replay_overlap(False)
    r.set_profiling(True)
    r.reset_stage_ms()
    for _ in range(20):
        r.replay()
    ms, runs = r.stage_ms()
    print(os.path.basename(api.LIB_PATH), f"band 1/{world} rows {y0}-{y1}:", {k: round(v / runs, 4) for k, v in ms.items()})
    r.close()
