"""CPU test of bench.py's reference arm (the one leg of the bench that needs no GPU): it runs, prints ONE JSON
line with the contract's keys, and its `config` object is exactly the one the B200 arm prints for the same
workload and N (the driver compares the two)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "shaded_gpixels_per_s" and d["unit"] == "Gpixels/s"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["cpu"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.make_config("views1080_tex", bench.WORKLOADS["views1080_tex"]["views"], 1)
    assert "configs[4]" in d["config"]["workload"] and d["config"]["views_per_step_per_gpu"] == 512


def test_other_ranks_of_the_reference_arm_exit_quietly(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_raster_stage_split_labels():
    """roofline.raster_kernels_ms_per_step names the kernels the stage really ran: the one-kernel opaque stage,
    the visibility + resolve pair, or the one raster kernel of a pass with blended primitives."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.kernel_split((4.0, 0.0), 2, 2, True) == {"raster_opaque_kernel<true>": 2.0}
    assert bench.kernel_split((4.0, 2.0), 2, 1, True) == {"raster_opaque_kernel<false>": 2.0, "resolve_kernel": 1.0}
    assert bench.kernel_split((4.0, 0.0), 2, 0, True) == {"raster_tex_kernel": 2.0}
    assert bench.kernel_split((4.0, 0.0), 0, 0, False) == {"raster_kernel": 4.0}
    assert bench.stage_label(2, True).startswith("raster_opaque_kernel<true>") and bench.stage_label(0, False) == "raster_kernel"


def test_gpu_arm_fails_loudly_without_a_device(built):
    """No CPU fallback: on a box without a GPU the B200 arm must not print a line."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline",
                          "--no-others"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
