"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic: view partition and the
sort-first band split + gather.  Each rank plays a GPU: it renders the frame with the CPU checker,
keeps ONLY its band (what a band-restricted context produces -- tests/test_gpu_parity.py proves
that equivalence on the device), and the gather must reassemble the full frame on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dtrenderer_b200 import multigpu, scenes


def test_split_views_partitions_exactly():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            got = sum((multigpu.split_views(n, world, r) for r in range(world)), [])
            assert got == list(range(n))


def test_band_rows_tile_aligned_and_exact():
    for h in (1, 31, 32, 33, 600, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = [multigpu.band_rows(h, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            for (a0, a1), (b0, b1) in zip(rows[:-1], rows[1:]):
                assert a1 == b0
            for y0, y1 in rows:
                assert y0 == y1 or (y0 % 32 == 0 and (y1 % 32 == 0 or y1 == h) and y0 < y1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dtro
    o = dtro.Oracle(w, h, "port")
    scenes.replay(scenes.cfg1_scene(w, h) + scenes.fill_scene(w, h, 3000)[1:], o)
    y0, y1 = multigpu.band_rows(h, world, rank)
    color = torch.full((h, w), -1, dtype=torch.int32)
    depth = torch.full((h, w), float("nan"), dtype=torch.float32)
    color[y0:y1] = torch.from_numpy(o.color().view(np.int32).copy())[y0:y1]
    depth[y0:y1] = torch.from_numpy(o.zbuffer().copy())[y0:y1]
    nbytes = multigpu.gather_bands(color, depth, h, dst=0)
    dist.barrier()
    if rank == 0:
        ok_c = np.array_equal(color.numpy().view(np.uint32), o.color())
        ok_z = np.array_equal(depth.numpy().view(np.uint32), o.zbuffer().view(np.uint32))
        expect = sum(8 * w * (b - a) for a, b in (multigpu.band_rows(h, world, r) for r in range(1, world)))
        out.put((ok_c, ok_z, nbytes == expect))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_gather_reassembles_the_frame(built, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 320, 200, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == (True, True, True)
