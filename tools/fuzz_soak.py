"""Soak run of the mixed-primitive fuzz (tests/test_gpu_parity.py::test_fuzz_mixed_primitives) over many
seeds and frame sizes on the GPU box: `python tools/fuzz_soak.py 60`.  Prints the failing seeds, if any.
`python tools/fuzz_soak.py 150 opaque`: only the all-opaque variant (one-kernel and two-kernel opaque stage)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import test_gpu_parity as t  # noqa: E402
import test_gpu_round2 as t2  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
only_opaque = len(sys.argv) > 2 and sys.argv[2] == "opaque"
rng = np.random.default_rng(77)
bad = []
for seed in range(100, 100 + n):
    size = (int(rng.integers(3, 1400)), int(rng.integers(3, 900)))
    try:
        if not only_opaque:
            t.test_fuzz_mixed_primitives(True, seed, size)
    except AssertionError as e:
        bad.append((seed, size, str(e)[:120]))
        print("FAIL", seed, size, str(e)[:200], flush=True)
    try:  # the all-opaque variant: the same calls through the one-kernel and the two-kernel opaque stage
        t2.test_fuzz_opaque_passes_take_the_deferred_stage(True, seed, size)
    except AssertionError as e:
        bad.append(("opaque", seed, size, str(e)[:120]))
        print("FAIL opaque", seed, size, str(e)[:200], flush=True)
print(f"fuzz soak: {n} seeds x {1 if only_opaque else 2} kinds, {len(bad)} failures", bad)
sys.exit(1 if bad else 0)
