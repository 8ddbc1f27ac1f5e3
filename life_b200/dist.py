"""Host-side plumbing for the slab-decomposed lattice: one process per GPU, launched by torchrun.

The data path between GPUs lives in liblife_b200 (csrc/halo.cu: paired ncclSend/ncclRecv of the 3 outgoing populations of one
column per face, overlapped with the interior sweep).  What remains for the host is bookkeeping, done here over
torch.distributed — with the nccl backend on the GPU box and with gloo in the CPU tests (tests/test_dist_gloo.py):

  * rendezvous of the ncclUniqueId liblife_b200 needs for its own communicator,
  * cutting the reference's global arrays (x-major, so a slab is one contiguous chunk; src/Grid.cpp:70) into per-rank
    slabs and putting per-rank results back together,
  * reductions of host scalars (max-over-ranks timing).

No lattice arithmetic happens here.
"""
import os

import numpy as np

from . import capi


def env_ranks():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when run plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def _dist():
    import torch.distributed as dist
    return dist


def is_parallel():
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def share_nccl_id(make_id=capi.nccl_unique_id):
    """Rank 0 creates the 128-byte ncclUniqueId, every rank returns the same bytes.  None when there is one rank."""
    if not is_parallel():
        return None
    dist = _dist()
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def slab_of(global_array, Nx, rank=None, world=None):
    """The chunk of a reference-layout global array (leading dimension Nx, or flat with Nx*k entries) owned by `rank`."""
    dist = _dist()
    if world is None:
        world = dist.get_world_size() if is_parallel() else 1
    if rank is None:
        rank = dist.get_rank() if is_parallel() else 0
    b, e = capi.slab_range(Nx, world, rank)
    a = np.asarray(global_array)
    if a.shape[0] != Nx:
        a = a.reshape(Nx, -1)
    return np.ascontiguousarray(a[b:e])


def scatter_slabs(global_array, Nx, src=0):
    """`global_array` (only needed on `src`) -> this rank's slab.  Point-to-point, one message per rank."""
    if not is_parallel():
        return slab_of(global_array, Nx, 0, 1)
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [None]
    if rank == src:
        a = np.asarray(global_array)
        meta[0] = (a.shape[1:], str(a.dtype))
    dist.broadcast_object_list(meta, src=src)
    tail, dtype = meta[0]
    b, e = capi.slab_range(Nx, world, rank)
    if rank == src:
        reqs = []
        for r in range(world):
            if r == src:
                continue
            rb, re_ = capi.slab_range(Nx, world, r)
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[rb:re_])), dst=r))
        for q in reqs:
            q.wait()
        return np.ascontiguousarray(a[b:e])
    out = torch.empty((e - b,) + tuple(tail), dtype=getattr(torch, dtype))
    dist.recv(out, src=src)
    return out.numpy()


def gather_slabs(local_array, Nx, dst=0):
    """Per-rank slabs -> the global array on `dst` (None elsewhere)."""
    if not is_parallel():
        return np.asarray(local_array)
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    loc = np.ascontiguousarray(local_array)
    if rank != dst:
        dist.send(torch.from_numpy(loc), dst=dst)
        return None
    out = np.empty((Nx,) + loc.shape[1:], dtype=loc.dtype)
    for r in range(world):
        rb, re_ = capi.slab_range(Nx, world, r)
        if r == dst:
            out[rb:re_] = loc
        else:
            buf = torch.empty((re_ - rb,) + loc.shape[1:], dtype=torch.from_numpy(loc).dtype)
            dist.recv(buf, src=r)
            out[rb:re_] = buf.numpy()
    return out


def max_over_ranks(x, device=None):
    """Maximum of a host scalar over all ranks (the time of a step is the slowest rank's)."""
    if not is_parallel():
        return float(x)
    import torch
    dist = _dist()
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else None)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def halo_plan(Nx, Ny, world, periodic_x):
    """What crosses each slab face per step, as (src_rank, dst_rank, populations, n_doubles) — the design csrc/halo.cu
    implements, stated on the host so it can be checked against the reference's push map without a GPU."""
    plan = []
    for r in range(world):
        right = r + 1 if r + 1 < world else (0 if periodic_x else None)
        left = r - 1 if r > 0 else (world - 1 if periodic_x else None)
        if right is not None and world > 1:
            plan.append((r, right, (1, 5, 7), 3 * Ny))     # cx = +1 leave through the right face
        if left is not None and world > 1:
            plan.append((r, left, (2, 6, 8), 3 * Ny))      # cx = -1 leave through the left face
    return plan


def file_plan(kind, Nx, Ny, world, rank, head_len=0, tail_len=0):
    """Which bytes of the ONE shared fluid file rank `rank` writes (the design csrc/lbm_file.cu implements with pwrite(), stated on
    the host so it can be checked without a GPU): a list of (file_offset, n_bytes, what), `what` naming the source:

      kind "restart" (src/Grid.cpp:1163-1229): ("head",) 44 bytes by rank 0; ("records", i_begin, i_end) = the slab's 120-byte
          records, one contiguous run because the file is i-major
      kind "vti" (src/Grid.cpp:790-898): ("head",), ("size", block) x 3 and ("tail",) by rank 0; ("row", block, j) = this slab's
          segment of row j of block 0 / 1 / 2 (Density, Pressure: 8 bytes per node; Velocity: 24), because the blocks are j-major
    """
    b, e = capi.slab_range(Nx, world, rank)
    plan = []
    if kind == "restart":
        if rank == 0:
            plan.append((0, 44, ("head",)))
        plan.append((44 + b * Ny * 120, (e - b) * Ny * 120, ("records", b, e)))
        return plan
    if kind != "vti":
        raise ValueError(kind)
    n8 = Nx * Ny * 8
    data = [head_len + 8, head_len + 8 + n8 + 8, head_len + 8 + 2 * (n8 + 8)]
    if rank == 0:
        plan.append((0, head_len, ("head",)))
        for blk in range(3):
            plan.append((data[blk] - 8, 8, ("size", blk)))
        plan.append((data[2] + 3 * n8, tail_len, ("tail",)))
    for blk in range(3):
        comp8 = 24 if blk == 2 else 8
        for j in range(Ny):
            plan.append((data[blk] + (j * Nx + b) * comp8, (e - b) * comp8, ("row", blk, j)))
    return plan
