"""Host-side plumbing for the two multi-GPU modes of the draw path.  One process per
GPU; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) carries the only exchange.

* frame / viewpoint parallel: view i -> rank i mod N; no collective on the data path.
* sort-first screen bands: rank r rasterises tile-aligned rows [y0, y1) of the SAME frame (every rank
  runs setup over all primitives but bins and rasterises only its band).  Two ways to assemble the
  frame on rank 0:
    - share_frames(): rank 0 exports CUDA IPC handles of its planes once; the other ranks map them
      and their raster kernels write finished regions straight into rank 0's HBM over NVLink (the
      transfer is fused into the kernel's write-back, tile by tile; a step ends with a stream-ordered
      barrier, nothing is copied afterwards);
    - gather_bands(): one grouped NCCL send/recv per frame lands every band in rank 0's planes (the
      baseline the peer-write path is measured against; also what the gloo CPU tests exercise).
"""
import numpy as np

TILE_H = 32  # default of dtr::TILE_H; callers with a loaded library pass dtr_b200_tile_height() instead


def split_views(n_views, world, rank):
    """Contiguous block partition of view indices: the views rank `rank` renders."""
    base, rem = divmod(n_views, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def band_rows(height, world, rank, tile_h=None):
    """Tile-aligned row band [y0, y1) of rank `rank`; bands tile [0, height) exactly, and trailing
    ranks may get an empty band when there are fewer tile rows than ranks.  The same partition as
    dtr_b200_band_rows (tests/test_capi_boundary.py compares them); tile_h = dtr_b200_tile_height()."""
    tile_h = tile_h or TILE_H
    tiles = (height + tile_h - 1) // tile_h
    base, rem = divmod(tiles, world)
    t0 = rank * base + min(rank, rem)
    t1 = t0 + base + (1 if rank < rem else 0)
    return min(t0 * tile_h, height), min(t1 * tile_h, height)


class _DevicePtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def frame_tensors(renderer, frame=0):
    """torch views (int32 colour, float32 depth) of a frame's planes in HBM -- no copy."""
    import torch
    c, z = renderer.frame_device_ptrs(frame)
    shape = (renderer.height, renderer.width)
    col = torch.as_tensor(_DevicePtr(c, shape, "<i4"), device="cuda")
    dep = torch.as_tensor(_DevicePtr(z, shape, "<f4"), device="cuda")
    return col, dep


def gather_bands(color, depth, height, dst=0, group=None):
    """Every rank's band rows of `color`/`depth` ([H, W] tensors, same shape on all ranks) are sent
    into the same rows of rank `dst`'s tensors.  Rows are contiguous, so each band is one message per
    plane and lands in place."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ops = []
    if rank == dst:
        for r in range(world):
            y0, y1 = band_rows(height, world, r)
            if r == dst or y1 <= y0:
                continue
            ops.append(dist.P2POp(dist.irecv, color[y0:y1], r, group))
            ops.append(dist.P2POp(dist.irecv, depth[y0:y1], r, group))
    else:
        y0, y1 = band_rows(height, world, rank)
        if y1 > y0:
            ops.append(dist.P2POp(dist.isend, color[y0:y1], dst, group))
            ops.append(dist.P2POp(dist.isend, depth[y0:y1], dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return int(sum(int(np.prod(o.tensor.shape)) * o.tensor.element_size() for o in ops))


def share_frames(renderer, dst=0, group=None):
    """Sort-first bands over peer memory: rank `dst` exports its frame planes, every other rank maps
    them and renders into them from now on.  Collective (one 128-byte broadcast)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == dst:
        hc, hz = renderer.export_frames()
        buf.copy_(torch.tensor(list(hc + hz), dtype=torch.uint8))
    dist.broadcast(buf, src=dst, group=group)
    if rank != dst:
        raw = bytes(buf.cpu().tolist())
        renderer.open_peer_frames(raw[:64], raw[64:])


def band_barrier(token, group=None):
    """Stream-ordered barrier after a band-split step: when it completes on rank 0 every rank's
    raster kernel (enqueued before it on the same stream) has finished writing its band."""
    import torch.distributed as dist
    dist.all_reduce(token, group=group)
