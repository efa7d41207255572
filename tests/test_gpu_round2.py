"""GPU parity (-m gpu), round 2 additions: the band exchange behind the C ABI (NCCL, two processes on
two GPUs), the device-side DTRMesh flattening, in-place asset updates, the texture edge uv == 1, the
frame-start ordering fixes, and the inexact-triangle replay at full 4K size."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from dtrenderer_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle(w, h):
    from oracle import dtro
    return dtro.Oracle(w, h, "reference" if dtro.available("reference") else "port")


def _renderer(w, h, frames=1, device=0):
    from dtrenderer_b200 import api
    return api.Renderer(w, h, frames, device)


def _same(col, z, o):
    assert np.array_equal(z.view(np.uint32), o.zbuffer().view(np.uint32)), "depth differs"
    assert np.array_equal(col, o.color()), "colour differs"


def test_uv_exactly_one_stays_inside_the_texture_allocation(built):
    """v == 1 indexes texel row h (DTRendererRender.cpp:1196-1203).  The reference build asserts there
    (`texelYf < dim.h`, DTR_DEBUG 1) and would otherwise read whatever follows its bitmap, so this case
    is checked against the C restatement, which has no assert: it is handed a bitmap followed by zeros,
    which is what the device allocation's spare row holds, so the frames must agree bit for bit."""
    w, h, tw, th = 200, 160, 16, 8
    rng = np.random.default_rng(5)
    buf = np.zeros((th + 2, tw, 4), np.uint8)
    buf[:th] = scenes.random_texture(tw, th, 3, opaque=False)
    tex = buf[:th]  # contiguous view: rows th, th+1 of buf (zeros) follow it in memory
    from oracle import dtro
    o, r = dtro.Oracle(w, h, "port"), _renderer(w, h)
    r.begin_frame(0)
    for t in (o, r):
        t.clear((0.2, 0.3, 0.4))
    for _ in range(6):
        p = np.concatenate([rng.integers(0, [w, h], (3, 2)), rng.uniform(0, 9, (3, 1))], 1).astype(np.float32)
        uv = rng.choice([0.0, 1.0], (3, 2)).astype(np.float32)  # corners of the texture, including (1, 1)
        for t in (o, r):
            t.textured_triangle(p.reshape(-1), uv.reshape(-1), tex, (1, 1, 1, 1), scenes.DEFAULT_TRIANGLE_TRANSFORM)
    col, z = r.end_frame(0)
    _same(col, z, o)


def test_clear_recorded_before_begin_frame_is_kept(built):
    """DTRRender_Clear writes the colour buffer and the frame start only resets depth: a clear issued
    just before dtr_b200_begin_frame(frame, NULL, NULL) must survive it."""
    w, h = 96, 64
    o, r = _oracle(w, h), _renderer(w, h)
    p = np.array([5, 5, 1, 80, 10, 1, 40, 60, 1], np.float32)
    r.clear((0.9, 0.1, 0.4))
    r.begin_frame(0)
    r.triangle(p, (0.1, 0.8, 0.2, 0.5), scenes.DEFAULT_TRIANGLE_TRANSFORM)
    col, z = r.end_frame(0)
    o.clear((0.9, 0.1, 0.4))
    o.reset_z()
    o.triangle(p, (0.1, 0.8, 0.2, 0.5), scenes.DEFAULT_TRIANGLE_TRANSFORM)
    _same(col, z, o)
    # ... and an uploaded colour buffer replaces a clear that came before it
    host = np.full((h, w), 0x00123456, np.uint32)
    r.clear((0.0, 1.0, 0.0))
    r.begin_frame(0, color=host)
    col, _ = r.end_frame(0)
    assert np.array_equal(col, host)


def _arena_mesh(mesh, order_seed=0):
    """The mesh as DTRAsset_LoadWavefrontObj leaves it: MeshFace[nF] whose three index pointers lead
    into one memory block (DTRendererAsset.cpp:509-578); the per-face arrays are laid out in a shuffled
    order with gaps, so the device really has to chase the pointers."""
    from dtrenderer_b200 import api
    f9 = np.asarray(mesh["faces"], np.int32)
    nF = f9.shape[0]
    rng = np.random.default_rng(order_seed)
    slots = rng.permutation(3 * nF)
    arena = np.zeros(3 * nF * 5 + 7, np.int32)  # 5 words per array slot: 3 used, 2 of padding
    arena[:] = -12345
    faces = (api.MeshFace * nF)()
    base = arena.ctypes.data
    for i in range(nF):
        for a, name in enumerate(("vertexIndex", "texIndex", "normalIndex")):
            at = int(slots[3 * i + a]) * 5 + 1
            arena[at:at + 3] = f9[i, 3 * a:3 * a + 3]
            setattr(faces[i], name, base + 4 * at)
        faces[i].numVertexIndex = faces[i].numTexIndex = faces[i].numNormalIndex = 3
    return faces, arena.view(np.uint8)


def test_mesh_faces_flattened_on_the_device(built):
    from dtrenderer_b200 import api
    w, h = 480, 270
    mesh, tex = scenes.uv_sphere(20, 10), scenes.random_texture(64, 64, 2, True)
    tr = scenes.transform7(40.0, (0, 1, 0), (1, 1, 1))
    o, r = _oracle(w, h), _renderer(w, h)
    faces, arena = _arena_mesh(mesh)
    mid = r.upload_mesh_faces(mesh, faces, arena, r.upload_texture(tex))
    r.begin_frame(0)
    r.clear((0, 0, 0))
    r.mesh_id(mid, scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), (0, 0, 0), tr)
    col, z = r.end_frame(0)
    o.clear((0, 0, 0))
    o.mesh(mesh, tex, scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), (0, 0, 0), tr)
    _same(col, z, o)
    # what the reference asserts on is an argument error here: a pointer outside the block, a bad count,
    # an index out of range
    for breakage in ("pointer", "count", "index"):
        faces, arena = _arena_mesh(mesh)
        if breakage == "pointer":
            faces[7].texIndex = arena.ctypes.data + arena.nbytes + 64
        elif breakage == "count":
            faces[3].numNormalIndex = 4
        else:
            arena.view(np.int32)[(faces[11].vertexIndex - arena.ctypes.data) // 4] = 10 ** 6
        with pytest.raises(api.DtrError):
            r.upload_mesh_faces(mesh, faces, arena, -1)


def test_update_texture_in_place(built):
    w, h = 160, 120
    tex = scenes.random_texture(32, 32, 7, opaque=False)
    o, r = _oracle(w, h), _renderer(w, h)
    tid = r.upload_texture(tex)
    tex2 = scenes.random_texture(32, 32, 8, opaque=False)
    r.update_texture(tid, tex2)
    r.begin_frame(0)
    r.clear((0.5, 0.5, 0.5))
    r.bitmap_id(tid, (20, 10), scenes.transform7(0.2, (0.5, 0.5, 0.5), (3, 3, 1)))
    col, z = r.end_frame(0)
    o.clear((0.5, 0.5, 0.5))
    o.bitmap(tex2, (20, 10), scenes.transform7(0.2, (0.5, 0.5, 0.5), (3, 3, 1)))
    _same(col, z, o)


_WORKER = r"""
import os, sys, time, numpy as np
sys.path.insert(0, {root!r})
from dtrenderer_b200 import api, scenes
rank, world, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
w, h = 1024, 768
scene = scenes.fill_scene(w, h, 20000, seed=11) + scenes.cfg1_scene(w, h)[1:]

def publish(path, data):
    open(path + ".tmp", "wb").write(data)
    os.rename(path + ".tmp", path)

def fetch(path):
    while not os.path.exists(path): time.sleep(0.01)
    return open(path, "rb").read()

def make(tag):
    # one band context: its own NCCL communicator and (peer modes) its own mapping of rank 0's planes;
    # the unique id and the IPC handles travel through files
    r = api.Renderer(w, h, 1, rank)
    base = sys.argv[4] + tag
    if rank == 0:
        publish(base, r.band_comm_unique_id())
    r.band_comm_init(fetch(base), world, rank)
    r.set_band(*r.band_rows(world, rank))
    if mode != "nccl":
        if rank == 0:
            hc, hz = r.export_frames()
            publish(base + ".ipc", hc + hz)
        else:
            raw = fetch(base + ".ipc")
            r.open_peer_frames(raw[:64], raw[64:])
        r.band_barrier()  # rank 0's planes exist and are mapped before anybody writes
    return r

# "peer2": two frame targets per rank used alternately (what bench.py's band loop does): the barrier that
# completes frame i runs beside the rasterisation of frame i+1, each context on its own stream
ctxs = [make("a"), make("b")] if mode == "peer2" else [make("a")]
for rep in range(4):
    r = ctxs[rep % len(ctxs)]
    r.begin_frame(0)
    scenes.replay(scene, r)
    if mode == "nccl":
        r.gather_bands(0, 0)
    else:
        r.band_barrier()
if rank == 0:
    whole = api.Renderer(w, h, 1, 0)
    whole.begin_frame(0)
    scenes.replay(scene, whole)
    c1, z1 = whole.end_frame(0)
    ok = True
    for r in ctxs:
        col, z = r.end_frame(0)
        ok = ok and np.array_equal(col, c1) and np.array_equal(z.view(np.uint32), z1.view(np.uint32))
    print("ASSEMBLED_OK" if ok else "ASSEMBLED_DIFFERS", flush=True)
for r in ctxs:
    r.band_barrier()  # nobody tears its context down while another rank still uses it
    r.sync()
"""


@pytest.mark.parametrize("mode", ["nccl", "peer", "peer2"])
def test_band_exchange_through_the_c_abi_two_processes(built, tmp_path, mode):
    """Sort-first bands with NO Python on the data path: one process per GPU, the library's own NCCL
    communicator (dtr_b200_band_comm_init), dtr_b200_gather_bands (grouped ncclSend/ncclRecv) or the
    peer-memory write-back ended by dtr_b200_band_barrier -- with one frame target per rank or with two
    used alternately ("peer2": double buffering, each context on its own stream with its own
    communicator); rank 0's assembled frame(s) must equal its own whole-frame render bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    rendezvous = str(tmp_path / f"uid_{mode}")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    procs = [subprocess.Popen([sys.executable, str(script), str(rk), "2", mode, rendezvous], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=env) for rk in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "ASSEMBLED_OK" in outs[0], outs


def test_rotated_screen_size_triangles_at_4k(built):
    """Inexact triangles (user rotation / scale: non-integer vertices) replay the reference's sequential
    fp32 accumulation.  Screen-size ones at 4K used to cost O(bbox) adds per pixel; with the row-shared
    replay the frame below takes milliseconds, and it must still match the reference bit for bit."""
    import time
    w, h = 3840, 2160
    rng = np.random.default_rng(2024)
    o, r = _oracle(w, h), _renderer(w, h)
    r.begin_frame(0)
    for t in (o, r):
        t.clear((0.1, 0.1, 0.2))
    tris = []
    for i in range(5):
        p = np.concatenate([rng.uniform([-200, -200], [w + 200, h + 200], (3, 2)), rng.uniform(0, 255, (3, 1))], 1).astype(np.float32)
        col = np.concatenate([rng.random(3), [1.0 if i % 2 == 0 else 0.6]]).astype(np.float32)
        tr = scenes.transform7(float(rng.uniform(-1.5, 1.5)), (0.33, 0.33, 0.33), (float(rng.uniform(0.7, 1.3)), float(rng.uniform(0.7, 1.3)), 1.0))
        tris.append((p.reshape(-1), col, tr))
    # plus non-integer vertices without any rotation, and a sliver
    tris.append((np.array([10.5, 20.25, 5, 3800.75, 100.5, 200, 1900.125, 2100.5, 90], np.float32), np.array([0.2, 0.9, 0.3, 0.8], np.float32),
                 scenes.DEFAULT_TRIANGLE_TRANSFORM))
    tris.append((np.array([0.5, 0.5, 1, 3839.5, 2159.5, 250, 3839.5, 2150.25, 100], np.float32), np.array([0.9, 0.9, 0.1, 1.0], np.float32),
                 scenes.DEFAULT_TRIANGLE_TRANSFORM))
    for p, col, tr in tris:
        for t in (o, r):
            t.triangle(p, col, tr)
    t0 = time.perf_counter()
    col, z = r.end_frame(0)
    gpu_s = time.perf_counter() - t0
    _same(col, z, o)
    assert gpu_s < 2.0, f"the 4K replay frame took {gpu_s:.2f} s"


# ---- the deferred raster stage (visibility + resolve kernels) ----------------------------------------
_DEFER_SCRIPT = r"""
import hashlib, sys, numpy as np
sys.path.insert(0, %(root)r)
from dtrenderer_b200 import api, scenes
w, h = 517, 301   # not a multiple of the tile size or of 4: the scalar load/store paths
r = api.Renderer(w, h, 2, 0)
mesh, tex = scenes.uv_sphere(), scenes.random_texture(64, 32, 9, opaque=True)
rng = np.random.default_rng(11)
for f in range(2):
    r.begin_frame(f)
    r.clear((0.1, 0.6, 0.3))
    r.mesh(mesh, tex, scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), (0, 0, 0),
           scenes.transform7(20.0 + 50 * f, (0, 1, 0), (1, 1, 1)))
    for _ in range(40):   # rotated (inexact) and plain flat triangles over the mesh, opaque
        p = np.concatenate([rng.integers(-40, [w + 40, h + 40], (3, 2)), rng.uniform(0, 12, (3, 1))], 1).astype(np.float32)
        r.triangle(p.reshape(-1), (*rng.random(3).tolist(), 1.0),
                   scenes.transform7(float(rng.uniform(0, 90)), (0.33, 0.33, 0.33), (1, 1, 1)) if rng.random() < 0.5 else None)
r.flush()
d = r.last_pass_stage()
out = hashlib.sha256()
for f in range(2):
    col, z = r.end_frame(f)
    out.update(col.tobytes()); out.update(z.tobytes())
print("RESULT", int(d), out.hexdigest(), r.stats()["setPixels"])
"""


def _run_defer_script(**overrides):
    env = dict(os.environ)
    env.pop("DTR_B200_DEFER", None)
    env.pop("DTR_B200_FUSED", None)
    env.update(overrides)
    p = subprocess.run([sys.executable, "-c", _DEFER_SCRIPT % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    return int(line[1]), line[2], line[3]


def test_deferred_stage_equals_the_single_kernel_stage(built):
    """A pass made of opaque triangles runs as raster_opaque_kernel<true> (visibility + in-place resolve);
    DTR_B200_FUSED=0 makes it visibility kernel + resolve kernel, DTR_B200_DEFER=0 the single-kernel
    stage.  Same frames (colour and depth, both frames, odd frame size, exact and inexact triangles)
    and the same SetPixel count in all three."""
    d2, h2, n2 = _run_defer_script()
    d1, h1, n1 = _run_defer_script(DTR_B200_FUSED="0")
    d0, h0, n0 = _run_defer_script(DTR_B200_DEFER="0")
    assert (d2, d1, d0) == (2, 1, 0), "the opaque stage did not engage / did not switch"
    assert h2 == h0, "one-kernel opaque stage and single-kernel frames differ"
    assert h1 == h0, "two-kernel opaque stage and single-kernel frames differ"
    assert n2 == n0 and n1 == n0, "SetPixel counts differ"


def test_deferred_stage_only_for_passes_that_never_blend(built):
    """One translucent primitive, a texture with a non-opaque texel, or a frame that continues from
    its previous contents: the pass takes the single-kernel stage.  Every case is checked against the
    oracle as well."""
    w, h = 256, 160
    mesh = scenes.uv_sphere()
    tr = scenes.transform7(35.0, (0, 1, 0), (1, 1, 1))
    opaque_tex = scenes.random_texture(32, 32, 4, opaque=True)
    alpha_tex = scenes.random_texture(32, 32, 4, opaque=False)

    def run(tex, rect_alpha=None, second_pass=False):
        o, r = _oracle(w, h), _renderer(w, h)
        r.begin_frame(0)
        for t in (o, r):
            t.clear((0.3, 0.3, 0.7))
            t.mesh(mesh, tex, scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), (0, 0, 0), tr)
            if rect_alpha is not None:
                t.rectangle(mn=(10.0, 20.0), mx=(90.0, 70.0), color=(1.0, 0.5, 0.2, rect_alpha), transform=scenes.DEFAULT_TRANSFORM)
        r.flush()
        first = r.last_pass_deferred()
        second = None
        if second_pass:  # more opaque triangles onto the frame as it is: no clear, so nothing to defer onto
            p = np.array([20, 20, 3, 200, 40, 3, 90, 150, 3], np.float32)
            for t in (o, r):
                t.triangle(p, (0.9, 0.1, 0.1, 1.0), scenes.DEFAULT_TRIANGLE_TRANSFORM)
            r.flush()
            second = r.last_pass_deferred()
        col, z = r.end_frame(0)
        _same(col, z, o)
        return first, second

    assert run(opaque_tex) == (True, None)
    assert run(alpha_tex)[0] is False
    assert run(opaque_tex, rect_alpha=0.5)[0] is False
    assert run(opaque_tex, second_pass=True) == (True, False)


@pytest.mark.parametrize("seed,size", [(1, (257, 131)), (2, (640, 360)), (3, (97, 1000)), (4, (1023, 66)), (5, (1920, 1080))])
def test_fuzz_opaque_passes_take_the_deferred_stage(built, seed, size):
    """Random passes made ONLY of opaque triangles -- integer-vertex batches, transformed (inexact)
    triangles, textured triangles and meshes in all three light modes with an opaque texture -- in one
    flush at awkward sizes: the deferred stage engages, and colour, depth and the SetPixels counter
    equal the oracle's (the counter is order dependent: every fragment that passes the depth test when
    it is submitted counts, whether or not a later one replaces it)."""
    from dtrenderer_b200 import api
    rng = np.random.default_rng(2000 + seed)
    w, h = size
    tex = scenes.random_texture(int(rng.integers(2, 64)), int(rng.integers(2, 64)), 30 + seed, opaque=True)
    mesh = scenes.uv_sphere(14, 7)
    o, r, r2 = _oracle(w, h), _renderer(w, h), _renderer(w, h)
    r2.set_opaque_stage(api.OPAQUE_TWO_KERNELS)  # the same calls through the visibility + resolve pair
    o.reset_counters()
    r.begin_frame(0)
    r2.begin_frame(0)
    both = (o, r, r2)
    clear = tuple(rng.random(3))
    for t in both:
        t.clear(clear)
    for i in range(90):
        kind = int(rng.integers(0, 4))
        col = (*rng.random(3).tolist(), 1.0)
        if kind == 0:
            n = int(rng.integers(1, 60))
            c = rng.integers(-10, [w + 10, h + 10], (n, 1, 2))
            p = np.concatenate([c + rng.integers(-40, 41, (n, 3, 2)), rng.uniform(0, 255, (n, 3, 1))], 2).astype(np.float32)
            cols = rng.random((n, 4)).astype(np.float32)
            cols[:, 3] = 1.0
            for t in both:
                t.triangles(p.reshape(n, 9), cols, scenes.DEFAULT_TRIANGLE_TRANSFORM)
        elif kind == 1:
            p = np.concatenate([rng.uniform(-40, [w + 40, h + 40], (3, 2)), rng.uniform(0, 255, (3, 1))], 1).astype(np.float32)
            tri_tr = scenes.transform7(float(rng.uniform(-3, 3)), tuple(rng.random(3)),
                                       (float(rng.uniform(0.4, 2.5)), float(rng.uniform(0.4, 2.5)), 1.0))
            for t in both:
                t.triangle(p.reshape(-1), col, tri_tr)
        elif kind == 2:
            p = np.concatenate([rng.integers(-20, [w + 20, h + 20], (3, 2)), rng.uniform(0, 255, (3, 1))], 1).astype(np.float32)
            uv = (rng.random(6) * 0.999).astype(np.float32)
            for t in both:
                t.textured_triangle(p.reshape(-1), uv, tex, col, scenes.DEFAULT_TRIANGLE_TRANSFORM)
        elif i % 4 == 0:
            mode = int(rng.integers(0, 3))
            view = scenes.view_transforms(4096)[int(rng.integers(0, 4096))]
            mtex = tex if rng.random() < 0.5 else scenes.WHITE_TEXTURE
            pos = (float(rng.uniform(-0.3, 0.3)), 0.0, 0.0)
            for t in both:
                t.mesh(mesh, mtex, mode, (1, -1, 1), (1, 1, 1, 1), pos, view)
    for t, stage in ((r, api.OPAQUE_ONE_KERNEL), (r2, api.OPAQUE_TWO_KERNELS)):
        t.flush()
        assert t.last_pass_stage() == stage, "an all-opaque pass onto cleared frames did not take the deferred stage"
        col_, z_ = t.end_frame(0)
        _same(col_, z_, o)
        assert t.stats()["setPixels"] == o.counters()[0]


@pytest.mark.parametrize("stage", [2, 1])
@pytest.mark.parametrize("two_gpus", [False, True])
def test_deferred_stage_with_foreign_output_planes(built, two_gpus, stage):
    """A band context whose output planes belong to ANOTHER context (the peer write-back of the band
    split).  One-kernel stage: finished regions go straight to the output.  Two-kernel stage: the
    visibility kernel leaves its tags in the band context's own colour planes, the resolve kernel
    stores every pixel of the busy tiles to the output.  Two bands of an all-opaque scene (odd frame
    width: the scalar paths), assembled in the first context's planes, equal the oracle's frame."""
    import torch
    if two_gpus and torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dtrenderer_b200 import api
    w, h = 1021, 600
    tex = scenes.random_texture(32, 16, 8, opaque=True)
    rng = np.random.default_rng(21)
    n = 3000
    c = rng.integers(0, [w, h], (n, 1, 2))
    p = np.concatenate([c + rng.integers(-25, 26, (n, 3, 2)), rng.uniform(0, 255, (n, 3, 1))], 2).astype(np.float32)
    cols = rng.random((n, 4)).astype(np.float32)
    cols[:, 3] = 1.0
    rot = scenes.transform7(17.0, (0.33, 0.33, 0.33), (1, 1, 1))
    big = np.array([40, 30, 2, 980, 200, 2, 500, 590, 2], np.float32)

    def draw(t):
        t.clear((0.7, 0.2, 0.1))
        t.mesh(scenes.uv_sphere(), tex, scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), (0, 0, 0),
               scenes.transform7(40.0, (0, 1, 0), (1, 1, 1)))
        t.triangles(p.reshape(n, 9), cols, scenes.DEFAULT_TRIANGLE_TRANSFORM)
        t.triangle(big, (0.1, 0.9, 0.4, 1.0), rot)  # inexact: shaded inside the visibility kernel

    o = _oracle(w, h)
    o.reset_counters()
    draw(o)
    r0, r1 = api.Renderer(w, h, 1, 0), api.Renderer(w, h, 1, 1 if two_gpus else 0)
    if two_gpus:
        r1.enable_peer_access(0)
    c0, z0 = r0.frame_device_ptrs(0)
    r1.set_output_planes(c0, z0)
    y0, y1 = r0.band_rows(2, 0)
    r0.set_band(y0, y1)
    r1.set_band(*r1.band_rows(2, 1))
    for r in (r0, r1):
        r.set_opaque_stage(stage)
        r.begin_frame(0)
        draw(r)
        r.flush()
        assert r.last_pass_stage() == stage
    r1.sync()
    col, z = r0.end_frame(0)
    _same(col, z, o)
    assert r0.stats()["setPixels"] + r1.stats()["setPixels"] == o.counters()[0]


def test_raster_stage_split_timing(built):
    """dtr_b200_get_raster_split_ms: with profiling on, the two-kernel opaque stage reports both of its
    kernels and their sum is the raster entry of dtr_b200_get_stage_ms; the one-kernel opaque stage and a
    blended pass report a single kernel."""
    from dtrenderer_b200 import api
    w, h = 640, 360
    mesh, tex = scenes.uv_sphere(), scenes.random_texture(32, 32, 4, opaque=True)
    tr = scenes.transform7(35.0, (0, 1, 0), (1, 1, 1))
    for blended, stage in ((False, api.OPAQUE_ONE_KERNEL), (False, api.OPAQUE_TWO_KERNELS), (True, api.OPAQUE_ONE_KERNEL)):
        r = _renderer(w, h)
        r.set_opaque_stage(stage)
        r.set_profiling(True)
        r.begin_frame(0)
        r.clear((0.3, 0.3, 0.7))
        r.mesh(mesh, tex, scenes.SHADE_GOURAUD, (1, -1, 1), (1, 1, 1, 1), (0, 0, 0), tr)
        if blended:
            r.rectangle(mn=(10.0, 20.0), mx=(90.0, 70.0), color=(1.0, 0.5, 0.2, 0.5), transform=scenes.DEFAULT_TRANSFORM)
        r.flush()
        for _ in range(3):
            r.replay()
        stage_ms, runs = r.stage_ms()
        (first, second), runs2 = r.raster_split_ms()
        assert runs == runs2 == 4
        assert r.last_pass_stage() == (0 if blended else stage)
        assert first > 0.0
        if blended or stage == api.OPAQUE_ONE_KERNEL:
            assert second < 0.02 * runs  # an empty interval between two events: a few microseconds per run
        else:
            assert second > 0.0
        assert abs((first + second) - stage_ms["raster"]) <= 0.05 * stage_ms["raster"] + 0.01
        r.close()
