"""Details of a deferred-stage fuzz mismatch: python tools/debug_deferred.py [seed w h]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_round2 as t2

def same(col, z, o):
    oc, oz = o.color(), o.zbuffer()
    dc = np.argwhere(col != oc)
    dz = np.argwhere(z.view(np.uint32) != oz.view(np.uint32))
    print(f"colour diffs {len(dc)} depth diffs {len(dz)}")
    for y, x in dc[:12]:
        print("   ", (int(x), int(y)), hex(int(col[y, x])), "vs", hex(int(oc[y, x])), "z", z[y, x])
t2._same = same
seed, w, h = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 257, 131)
try:
    t2.test_fuzz_opaque_passes_take_the_deferred_stage(True, seed, (w, h))
except AssertionError as e:
    print("assert", e)
