"""Where a band step's time goes on ONE GPU: a 1/N band of configs[3] replayed K times with the overlap on;
device time per step (CUDA events) against the host's enqueue time per step (wall clock before the sync),
with and without stage profiling.  usage: python tools/band_probe2.py [N ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dtrenderer_b200 import api, multigpu, scenes  # noqa: E402

w, h, n = 3840, 2160, 1_000_000
p, color = scenes.small_triangles(w, h, n, seed=7)
torch.cuda.init()
for world in [int(a) for a in sys.argv[1:]] or [1, 4, 8]:
    r = api.Renderer(w, h, 1, 0)
    y0, y1 = multigpu.band_rows(h, world, 0, r.tile_height())
    if world > 1:
        r.set_band(y0, y1)
    r.begin_frame(0)
    r.clear((0, 0, 0))
    r.triangles(p, color, scenes.DEFAULT_TRIANGLE_TRANSFORM)
    r.flush()
    for prof in (False, True):
        r.set_profiling(prof)
        for _ in range(10):
            r.replay()
        r.sync()
        K = 200
        t0 = time.perf_counter()
        for _ in range(K):
            r.replay()
        t1 = time.perf_counter()
        r.sync()
        t2 = time.perf_counter()
        print(f"band 1/{world} profiling {prof}: host enqueue {1e3 * (t1 - t0) / K:.4f} ms/step, "
              f"wall incl. sync {1e3 * (t2 - t0) / K:.4f} ms/step", flush=True)
    r.close()
