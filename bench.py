#!/usr/bin/env python
"""bench.py -- throughput of the B200 draw path on BASELINE.json's headline workload.

    python bench.py --gpus N --steps K --warmup W            # this back end
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (N=1 and frame-parallel N>1): BASELINE.json configs[1] -- the ~2.5k-triangle synthetic
mesh, perspective, Gouraud, 1920x1080 colour+z.  One STEP renders a batch of B viewpoints of that
scene (B frames, each into its own colour+z target) with ONE pass of the pipeline
setup -> scan -> bin -> raster.  `value` times dtr_b200_replay() (command list and mesh resident in
HBM); `e2e` times the public call path with host buffers: record draw calls -> flush (H2D of the
command block) -> read the B colour frames back into pinned host memory.

Prints ONE JSON line (rank 0).  See DESIGN.md §5 for the definitions behind every field.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from dtrenderer_b200 import scenes  # noqa: E402

WORKLOADS = {
    # name: (width, height, textured, tex_size, description)
    "mesh1080": (1920, 1080, False, 1,
                 "BASELINE configs[1]: 2500-triangle UV-sphere mesh, perspective, Gouraud, 1920x1080 colour+z"),
    "mesh4k_tex": (3840, 2160, True, 1024,
                   "BASELINE configs[2]: same mesh, 1024^2 texture (nearest, gamma-2), 3840x2160 colour+z"),
    "views1080_tex": (1920, 1080, True, 1024,
                      "BASELINE configs[4]: textured mesh viewpoints at 1920x1080, frame-parallel"),
}
FILL = {"fill4k": (3840, 2160, 1_000_000,
                   "BASELINE configs[3]: fill-rate stress, 1M small random z-buffered triangles at 3840x2160; "
                   "N>1: sort-first screen bands gathered to rank 0 over NCCL")}
METRIC, UNIT = "shaded_gpixels_per_s", "Gpixels/s"
TRIS_PER_FRAME = 2500


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload):
    """dram bytes per raster launch from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "raster_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            pass
    return None


def view_args(first, n, total_views=4096):
    ts = scenes.view_transforms(total_views)
    sel = [ts[(first + i) % total_views] for i in range(n)]
    pos = np.zeros((n, 3), np.float32)
    return pos, sel


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own scalar path (oracle/_ref, unmodified sources) on host cores
# --------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(kind, w, h, textured, tex_size):
    from oracle import dtro
    _W["o"] = dtro.Oracle(w, h, kind)
    _W["mesh"] = scenes.uv_sphere()
    _W["tex"] = scenes.random_texture(tex_size, tex_size, 1, True) if textured else scenes.WHITE_TEXTURE
    _cpu_frame(0)  # warm caches / page in


def _cpu_frame(view):
    o = _W["o"]
    _, ts = view_args(view, 1)
    o.reset_z()
    o.reset_counters()
    t0 = time.perf_counter()
    o.clear((0.5, 0.0, 1.0))
    o.mesh(_W["mesh"], _W["tex"], scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), (0, 0, 0), ts[0])
    dt = time.perf_counter() - t0
    sp, tr = o.counters()
    return sp, tr, dt


class CpuPool:
    """One single-threaded reference renderer per host core (independent frames per core -- the
    deterministic way to use all cores)."""

    def __init__(self, workload, cores=None):
        import multiprocessing as mp
        from oracle import dtro
        self.kind = "reference" if dtro.available("reference") else "port"
        w, h, textured, tex_size, _ = WORKLOADS[workload]
        self.cores = cores or len(os.sched_getaffinity(0))
        self.pool = mp.get_context("fork").Pool(self.cores, _cpu_init, (self.kind, w, h, textured, tex_size))

    def run(self, first_view, n_frames):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_frame, range(first_view, first_view + n_frames), chunksize=1)
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), sum(r[1] for r in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(workload, target_cpu_seconds=15.0):
    pool = CpuPool(workload)
    sp1, _, w1 = pool.run(0, pool.cores)  # calibration pass = warm-up
    per_frame = max(w1, 1e-3)
    n = int(max(pool.cores, min(4096, target_cpu_seconds / per_frame * pool.cores)))
    n = (n // pool.cores) * pool.cores
    sp, tr, wall = pool.run(0, n)
    pool.close()
    return {"value": sp / wall / 1e9, "unit": UNIT, "cores": pool.cores, "kind": pool.kind,
            "sample": f"{n} frames (views 0..{n - 1}) of the workload, one single-threaded renderer per core, "
                      f"{wall:.2f} s wall", "frames_per_s": n / wall, "mtris_per_s": tr / wall / 1e6}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    w, h, _, _, desc = WORKLOADS[args.workload]
    pool = CpuPool(args.workload)
    # one step = the same batch of viewpoints as one step of the B200 arm (--views), spread over
    # all host cores; rounded up to whole rounds so that no core idles
    frames_per_step = max(1, -(-args.views // pool.cores)) * pool.cores
    for i in range(args.warmup):
        pool.run(i * frames_per_step, frames_per_step)
    sp = tr = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        a, b, _ = pool.run(i * frames_per_step, frames_per_step)
        sp += a
        tr += b
    wall = time.perf_counter() - t0
    pool.close()
    val = sp / wall / 1e9
    sample = (f"{frames_per_step} frames per step (the B200 arm's {args.views} viewpoints, one single-threaded "
              f"renderer per host core), {args.steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "width": w, "height": h, "triangles_per_frame": TRIS_PER_FRAME,
                   "frames_per_step": frames_per_step,
                   "parallelism": f"{pool.cores} host cores, independent frames per core"},
        "mtris_per_s": tr / wall / 1e6, "frames_per_s": frames_per_step * args.steps / wall,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": pool.cores, "kind": pool.kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread DURING the timed regions
# --------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:  # noqa: BLE001
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.002)

    def start(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def result(self):
        self._stop.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: keep this rank's host threads -- and with them the pinned readback buffers it
    allocates next -- on the CPUs closest to its GPU (NVML's ideal affinity), so that eight ranks'
    D2H streams do not all land on one socket's memory.  Best effort; a no-op if NVML says nothing."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_gpu_arm(args, rank, world, local_rank):
    w, h, textured, tex_size, desc = WORKLOADS[args.workload]
    B = args.views

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload)  # before CUDA is initialised in this process (fork safety)

    import torch
    import torch.distributed as dist
    from dtrenderer_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this back end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # 2*B frame targets: the device-resident measurement uses the first B; the end-to-end loop
    # alternates between the two halves so that a step's readback overlaps the next step's rendering
    r = api.Renderer(w, h, 2 * B, local_rank)
    # a real (non-NULL) stream: the kernels, the copies and the timing events all go on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    mesh = scenes.uv_sphere()
    tex = scenes.random_texture(tex_size, tex_size, 1, True) if textured else scenes.WHITE_TEXTURE
    first_view = rank * B
    pos, transforms = view_args(first_view, B)

    def record(first=0):
        for f in range(first, first + B):
            r.begin_frame(f)
            r.clear((0.5, 0.0, 1.0))
        r.mesh_views(mesh, tex, scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), pos, transforms, first)

    # one full pass to build the resident command list and count the work of a step
    r.reset_stats()
    record()
    r.flush()
    st = r.stats()
    shaded_per_step, tris_per_step = st["setPixels"], st["triangles"]
    upload_bytes = st["uploadBytes"]
    tex_bytes = 4 * min(shaded_per_step // B, tex_size * tex_size) if textured else 0
    alg_bytes = B * (8 * w * h + 156 * TRIS_PER_FRAME + tex_bytes)

    clocks = ClockSampler(local_rank)

    # ---- device-resident throughput: replay() ------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        r.replay()
    r.set_profiling(True)
    r.reset_stage_ms()
    r.reset_stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    e0.record(stream)
    for _ in range(args.steps):
        r.replay()
    e1.record(stream)
    barrier()
    clocks.pause()
    ms = e0.elapsed_time(e1)
    stage, runs = r.stage_ms()
    launches = r.stats()["kernelLaunches"]
    # untimed extra pass with the replay pipelining off: per-stage times in isolation (in the timed
    # region the pre-raster stages of replay i+1 overlap the raster kernel of replay i, so their
    # event intervals include the wait for free SMs and the stages no longer add up to the step)
    r.set_replay_overlap(False)
    r.reset_stage_ms()
    for _ in range(10):
        r.replay()
    stage_iso, runs_iso = r.stage_ms()
    r.set_replay_overlap(True)
    r.set_profiling(False)
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # ---- end to end: record -> flush (H2D) -> read B colour frames back (D2H, pinned) --------
    # Every step uploads its command block and reads its B colour planes back; the readback is
    # asynchronous (copy stream) and double buffered, so step i's D2H overlaps step i+1's work.
    host = torch.empty((2, B, h, w), dtype=torch.int32, pin_memory=True)
    e2e_steps = max(2, min(args.steps, args.e2e_steps))

    def time_e2e(read_half):
        def step_e2e(i):
            half = i & 1
            record(half * B)
            read_half(half)

        for i in range(2):
            step_e2e(i)
        r.wait_reads()
        barrier()
        clocks.start()
        e0.record(stream)
        for i in range(e2e_steps):
            step_e2e(i)
        r.wait_reads()  # blocks until the last copy has landed; e1 is recorded after that
        e1.record(stream)
        barrier()
        clocks.pause()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / e2e_steps

    e2e_ms = time_e2e(lambda half: r.read_frames_async_ptr(half * B, B, host[half].data_ptr()))
    # the same loop with the on-device 24-bit DIB encode (3 bytes per pixel over PCIe); reported beside
    # `e2e`, which stays the u32 DTRRenderBuffer layout
    pitch = r.bgr24_pitch()
    host24 = torch.empty((2, B, h, pitch), dtype=torch.uint8, pin_memory=True)
    e2e24_ms = time_e2e(lambda half: r.read_frames_bgr24_async_ptr(half * B, B, host24[half].data_ptr()))
    packed_ok = bool(torch.equal(host24[1, :, :, :3 * w].reshape(B, h, w, 3),
                                 host[1].view(torch.uint8).reshape(B, h, w, 4)[..., :3]))
    checksum = int(host[0, 0].view(-1)[::997].to(torch.int64).sum().item())

    if rank == 0:
        peak, peak_src = load_peaks()
        ms_step = ms_total / args.steps
        raster_ms = stage["raster"] / max(runs, 1)
        achieved = alg_bytes / (raster_ms * 1e-3) / 1e9 if raster_ms > 0 else 0.0
        traffic = load_traffic(args.workload)
        line = {
            "metric": METRIC, "value": world * shaded_per_step / (ms_step * 1e-3) / 1e9, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "triangles_per_frame": TRIS_PER_FRAME,
                       "views_per_step_per_gpu": B, "shaded_fragments_per_step_per_gpu": shaded_per_step,
                       "parallelism": f"frame-parallel x{world} (independent viewpoints per GPU, no collective)",
                       "l2": f"each step writes {B * 8 * w * h / 1e6:.0f} MB of frames per GPU (> 126 MB L2), "
                             "so no flush is needed between steps"},
            "mtris_per_s": world * tris_per_step / (ms_step * 1e-3) / 1e6,
            "frames_per_s": world * B / (ms_step * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "raster_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": raster_ms,
                         "stage_ms_per_step": {k: v / max(runs, 1) for k, v in stage.items()},
                         "stage_ms_isolated": {k: v / max(runs_iso, 1) for k, v in stage_iso.items()},
                         "note": "achieved/frac use the raster kernel's duration inside the timed region, where "
                                 "setup/scan/bin of the next replay run on a second stream during its tail (their "
                                 "stage_ms_per_step therefore include waiting for SMs); stage_ms_isolated = the same "
                                 "stages timed in an extra pass with that pipelining off"},
            "cpu_baseline": cpu,
            "e2e": {"value": world * shaded_per_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": upload_bytes, "d2h_bytes_per_step": B * 4 * w * h,
                    "ms_per_step": e2e_ms, "frames_per_s": world * B / (e2e_ms * 1e-3), "steps": e2e_steps,
                    "readback": "colour planes only (what the reference presents); depth stays in HBM; "
                                "asynchronous and double buffered (step i's D2H overlaps step i+1's rendering)",
                    "checksum": checksum},
            "e2e_bgr24": {"value": world * shaded_per_step / (e2e24_ms * 1e-3) / 1e9, "unit": UNIT,
                          "d2h_bytes_per_step": B * pitch * h, "ms_per_step": e2e24_ms,
                          "matches_u32_readback": packed_ok,
                          "readback": "dtr_b200_read_frames_bgr24_async: colour packed on the device into 24-bit "
                                      "bottom-up DIB rows (informational; `e2e` is the DTRRenderBuffer layout)"},
            "gpu_launches": launches, "clocks": clocks.result(),
        }
        print(json.dumps(line), flush=True)
    else:
        clocks.result()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# cfg 4: fill-rate stress, sort-first screen bands for N > 1
# --------------------------------------------------------------------------------------------
def _fill_cpu_init(kind, w, h):
    from oracle import dtro
    _W["o"] = dtro.Oracle(w, h, kind)


def _fill_cpu_chunk(args):
    p, color = args
    o = _W["o"]
    o.reset_z()
    o.reset_counters()
    o.clear((0, 0, 0))
    o.triangles(p, color, scenes.DEFAULT_TRIANGLE_TRANSFORM)
    return o.counters()


def run_fill_reference(args, rank, world):
    """Reference arm for cfg 4: every host core renders an independent frame of a bounded sample
    (n/cores... triangles of the same distribution) -- whole-frame order is inherently serial in the
    reference, so independent frames are the only deterministic way to use all cores."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import dtro
    w, h, n_full, desc = FILL[args.workload]
    kind = "reference" if dtro.available("reference") else "port"
    cores = len(os.sched_getaffinity(0))
    per_core = 20000
    chunks = []
    for i in range(cores):
        p, c = scenes.small_triangles(w, h, per_core, seed=100 + i)
        chunks.append((p, c))
    pool = mp.get_context("fork").Pool(cores, _fill_cpu_init, (kind, w, h))
    for _ in range(max(1, args.warmup)):
        pool.map(_fill_cpu_chunk, chunks, chunksize=1)
    sp = tr = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for a, b in pool.map(_fill_cpu_chunk, chunks, chunksize=1):
            sp += a
            tr += b
    wall = time.perf_counter() - t0
    pool.close()
    pool.join()
    val = sp / wall / 1e9
    sample = f"{cores} independent 4K frames of {per_core} triangles each per step (one per host core)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "width": w, "height": h, "triangles_per_step": per_core * cores},
        "mtris_per_s": tr / wall / 1e6,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}),
        flush=True)


def run_fill_gpu(args, rank, world, local_rank):
    w, h, n, desc = FILL[args.workload]
    if args.triangles:
        n = args.triangles
    import torch
    import torch.distributed as dist
    from dtrenderer_b200 import api, multigpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this back end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    r = api.Renderer(w, h, 1, local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    y0, y1 = multigpu.band_rows(h, world, rank)
    if world > 1:
        r.set_band(y0, max(y1, y0 + 1) if y1 > y0 else h)  # (empty bands cannot occur at 4K with N <= 8)
    p, color = scenes.small_triangles(w, h, n, seed=7)  # geometry replicated on every rank
    col_t, dep_t = multigpu.frame_tensors(r, 0)
    peer = world > 1 and args.gather == "peer"
    token = torch.zeros(1, device="cuda")
    if peer:
        multigpu.share_frames(r, dst=0)  # ranks > 0 now rasterise straight into rank 0's planes

    def record():
        r.begin_frame(0)
        r.clear((0, 0, 0))
        r.triangles(p, color, scenes.DEFAULT_TRIANGLE_TRANSFORM)

    r.reset_stats()
    record()
    r.flush()
    st = r.stats()
    shaded = torch.tensor([st["setPixels"]], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(shaded)
    shaded_per_step = int(shaded.item())
    upload_bytes = st["uploadBytes"]
    alg_bytes = 8 * w * (y1 - y0) + 156 * n  # this rank's band + all triangles (every rank runs setup)
    clocks = ClockSampler(local_rank)

    def step_resident():
        r.replay()
        if peer:
            multigpu.band_barrier(token)  # stream-ordered: every band has landed in rank 0's HBM
            return 0
        if world > 1:
            return multigpu.gather_bands(col_t, dep_t, h, dst=0)
        return 0

    for _ in range(max(args.warmup, 3)):
        step_resident()
    r.set_profiling(True)
    r.reset_stage_ms()
    r.reset_stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    e0.record(stream)
    nccl_bytes = 0
    for _ in range(args.steps):
        nccl_bytes = step_resident()
    e1.record(stream)
    barrier()
    clocks.pause()
    stage, runs = r.stage_ms()
    launches = r.stats()["kernelLaunches"]
    r.set_profiling(False)
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    host = torch.empty((h, w), dtype=torch.int32, pin_memory=True)
    e2e_steps = max(2, min(args.steps, args.e2e_steps))

    def step_e2e():
        record()
        r.flush()
        if peer:
            multigpu.band_barrier(token)
        elif world > 1:
            multigpu.gather_bands(col_t, dep_t, h, dst=0)
        if rank == 0:
            r.read_frames_ptr(0, 1, host.data_ptr())

    step_e2e()
    barrier()
    clocks.start()
    e0.record(stream)
    for _ in range(e2e_steps):
        step_e2e()
    e1.record(stream)
    barrier()
    clocks.pause()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps

    if rank == 0:
        peak, peak_src = load_peaks()
        raster_ms = stage["raster"] / max(runs, 1)
        achieved = alg_bytes / (raster_ms * 1e-3) / 1e9 if raster_ms > 0 else 0.0
        print(json.dumps({
            "metric": METRIC, "value": shaded_per_step / (ms_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": w, "height": h, "triangles_per_step": n,
                       "shaded_fragments_per_step": shaded_per_step,
                       "parallelism": (f"sort-first bands x{world}, " + ("bands written into rank 0's HBM over NVLink by the "
                                       "raster kernel's write-back (CUDA IPC peer memory), stream-ordered barrier"
                                       if peer else "grouped NCCL send/recv gather")) if world > 1
                       else "single GPU, whole frame",
                       "nccl_bytes_per_step_into_rank0": nccl_bytes,
                       "peer_bytes_per_step_into_rank0": (8 * w * (h - multigpu.band_rows(h, world, 0)[1])) if peer else 0,
                       "l2": "one 4K frame (66 MB) + 160 MB of primitive records per step; L2 is not flushed "
                             "between steps (frame planes fit in L2, records do not)"},
            "mtris_per_s": n / (ms_step * 1e-3) / 1e6, "frames_per_s": 1.0 / (ms_step * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "raster_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": load_traffic(args.workload), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": raster_ms,
                         "stage_ms_per_step": {k: v / max(runs, 1) for k, v in stage.items()},
                         "note": "rank 0's band; this config is ALU/ordering bound, not HBM bound"},
            "cpu_baseline": None,
            "e2e": {"value": shaded_per_step / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": upload_bytes,
                    "d2h_bytes_per_step": 4 * w * h, "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": launches, "clocks": clocks.result()}), flush=True)
    else:
        clocks.result()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mesh1080", choices=sorted(WORKLOADS) + sorted(FILL))
    ap.add_argument("--triangles", type=int, default=0, help="fill4k: override the triangle count")
    ap.add_argument("--views", type=int, default=64, help="viewpoints (frames) per step per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="fill4k, N>1: how the bands reach rank 0 (peer-memory write-back or NCCL send/recv)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in FILL:
        if args.impl == "reference":
            run_fill_reference(args, rank, world)
        else:
            run_fill_gpu(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
