"""Multi-rank host logic on CPU: torch.distributed with the gloo backend, world_size 2 and 3.

The GPU data path (csrc/halo.cu over NCCL) cannot run here; what can be checked without a GPU is everything around it:
  * the slab partition (life_slab_range) and the scatter / gather of reference-layout arrays (life_b200/dist.py),
  * the rendezvous of the ncclUniqueId,
  * the halo DESIGN: a numpy model of "push into a ghost ring, ship the three outgoing populations of one column per face
    to the neighbour, wrap y locally" — executed by real ranks exchanging real messages over gloo — lands every population
    exactly where the reference's push map recv = ((i+cx+Nx)%Nx)*Ny + (j+cy+Ny)%Ny (src/Grid.cpp:229,240) sends it.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CX = [0, 1, -1, 0, 0, 1, -1, 1, -1]
CY = [0, 0, 0, 1, -1, 1, -1, -1, 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Nx, Ny, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from life_b200 import capi, dist as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        # --- rendezvous of the communicator id (any 128 bytes will do on CPU)
        nid = D.share_nccl_id(make_id=lambda: bytes(range(128)))
        assert nid == bytes(range(128))

        # --- scatter / gather in the reference layout
        tags = np.arange(1, Nx * Ny * 9 + 1, dtype=np.float64).reshape(Nx, Ny, 9) if rank == 0 else None
        mine = D.scatter_slabs(tags, Nx)
        b, e = capi.slab_range(Nx, world, rank)
        assert mine.shape == (e - b, Ny, 9)
        assert mine[0, 0, 0] == b * Ny * 9 + 1
        back = D.gather_slabs(mine, Nx)
        if rank == 0:
            assert np.array_equal(back, tags)
        assert D.max_over_ranks(float(rank)) == world - 1

        # --- the halo design, executed: local push into a ghost ring, exchange per halo_plan, wrap y
        nxl = e - b
        ring = np.zeros((9, nxl + 2, Ny + 2))              # [v, c = il + 1, r = j + 1]
        for v in range(9):
            ring[v, 1 + CX[v]:1 + CX[v] + nxl, 1 + CY[v]:1 + CY[v] + Ny] = mine[:, :, v]
        # y wrap of every column incl. ghosts (periodic top/bottom: what k_wrap_y does before the exchange)
        for v in range(9):
            if CY[v] == 1:
                ring[v, :, 1] = ring[v, :, Ny + 1]
            elif CY[v] == -1:
                ring[v, :, Ny] = ring[v, :, 0]
        plan = D.halo_plan(Nx, Ny, world, periodic_x=True)
        sends = [p for p in plan if p[0] == rank]
        recvs = [p for p in plan if p[1] == rank]
        reqs, bufs = [], []
        for (_, dst, pops, n) in sends:
            col = nxl + 1 if pops == (1, 5, 7) else 0
            msg = torch.from_numpy(np.ascontiguousarray(np.stack([ring[v, col, 1:Ny + 1] for v in pops])))
            assert msg.numel() == n
            reqs.append(dist.isend(msg, dst=dst, tag=0 if pops == (1, 5, 7) else 1))
        for (src, _, pops, n) in recvs:
            buf = torch.empty((3, Ny), dtype=torch.float64)
            reqs.append(dist.irecv(buf, src=src, tag=0 if pops == (1, 5, 7) else 1))
            bufs.append((pops, buf))
        for q in reqs:
            q.wait()
        for pops, buf in bufs:
            col = 1 if pops == (1, 5, 7) else nxl            # cx=+1 arrive in my first column, cx=-1 in my last
            for k, v in enumerate(pops):
                ring[v, col, 1:Ny + 1] = buf[k].numpy()
        got = np.ascontiguousarray(np.transpose(ring[:, 1:nxl + 1, 1:Ny + 1], (1, 2, 0)))
        full = D.gather_slabs(got, Nx)
        if rank == 0:
            expect = np.zeros_like(tags)
            i, j = np.meshgrid(np.arange(Nx), np.arange(Ny), indexing="ij")
            for v in range(9):
                expect[(i + CX[v] + Nx) % Nx, (j + CY[v] + Ny) % Ny, v] = tags[i, j, v]
            assert np.array_equal(full, expect)
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (12, 7)), (3, (13, 5)), (2, (9, 16))])
def test_slab_plumbing_and_halo_design_over_gloo(world, shape, tmp_path, lib_built):
    import torch.multiprocessing as mp
    Nx, Ny = shape
    mp.spawn(_worker, args=(world, _free_port(), Nx, Ny, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def test_slab_partition_covers_the_lattice(lib_built):
    from life_b200 import capi
    for Nx in (4, 17, 501, 16384 * 8):
        for world in (1, 2, 3, 8):
            edges = [capi.slab_range(Nx, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == Nx
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            widths = [e - b for b, e in edges]
            assert max(widths) - min(widths) <= 1


def test_halo_plan_is_exactly_what_leaves_a_slab():
    """Against the oracle's restatement of the push map: the only (node, population) pairs whose target lies in another
    slab are the cx=+1 populations of a slab's last column and the cx=-1 populations of its first column."""
    from oracle import oracle as O
    from life_b200 import capi, dist as D
    Nx, Ny, world = 11, 6, 3
    o = O.Oracle(O.Params(Nx=Nx, Ny=Ny, wall_left=0, wall_right=0, wall_bottom=0, wall_top=0))
    owner = np.zeros(Nx, int)
    for r in range(world):
        b, e = capi.slab_range(Nx, world, r)
        owner[b:e] = r
    crossing = {}
    for i in range(Nx):
        for j in range(Ny):
            for v in range(9):
                ti = o.stream_target(i, j, v) // Ny
                if owner[ti] != owner[i]:
                    crossing.setdefault((owner[i], owner[ti]), set()).add((i, v))
    plan = D.halo_plan(Nx, Ny, world, periodic_x=True)
    assert {(s, d) for s, d, _, _ in plan} == set(crossing)
    for s, d, pops, n in plan:
        b, e = capi.slab_range(Nx, world, s)
        col = e - 1 if pops == (1, 5, 7) else b
        assert crossing[(s, d)] == {(col, v) for v in pops}
        assert n == 3 * Ny


# ---- the shared-file design of the device-fed file paths (csrc/lbm_file.cu with nranks > 1) ----------------------------------------
def _file_worker(rank, world, port, Nx, Ny, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from life_b200 import capi, dist as D
    from oracle import fluidfiles as F
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)                      # every rank builds the same global state and keeps its slab
        rho = 1.0 + 0.1 * rng.standard_normal((Nx, Ny))
        u = 0.05 * rng.standard_normal((Nx, Ny, 2))
        fi = 1e-3 * rng.standard_normal((Nx, Ny, 2))
        f = rng.random((Nx, Ny, 9))
        scal = dict(Dx=0.01, Dt=2e-4, Dm=1e-6, Drho=1.0)
        b, e = capi.slab_range(Nx, world, rank)
        head, tail = capi.vtk_frame(Nx, Ny, scal["Dx"])
        # what this rank would have in its staging buffers: the file image of ITS slab only
        my_vti = F.vti_bytes(rho[b:e], u[b:e], scal["Dx"], scal["Dt"], scal["Dm"], scal["Drho"], 1.0, 0.5)
        h_loc, _ = F.vti_frame(e - b, Ny, scal["Dx"])
        nl8 = (e - b) * Ny * 8
        blocks = [my_vti[len(h_loc) + 8:len(h_loc) + 8 + nl8], my_vti[len(h_loc) + 16 + nl8:len(h_loc) + 16 + 2 * nl8],
                  my_vti[len(h_loc) + 24 + 2 * nl8:len(h_loc) + 24 + 5 * nl8]]
        my_rst = F.restart_bytes(3, 1.7, scal["Dx"], scal["Dt"], scal["Dm"], rho[b:e], u[b:e], fi[b:e], f[b:e])
        # local record indices are slab-relative in that image: the device writes global i (k_restart_pack), so patch them
        rec = np.frombuffer(bytearray(my_rst[44:]), dtype=F._NODE).copy()
        rec["i"] += b
        for kind, name in (("vti", "Fluid.3.vti"), ("restart", "Fluid.restart")):
            path = os.path.join(out_dir, name)
            fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o666)
            for off, n, what in D.file_plan(kind, Nx, Ny, world, rank, len(head), len(tail)):
                if what[0] == "head":
                    data = head if kind == "vti" else F.restart_bytes(3, 1.7, scal["Dx"], scal["Dt"], scal["Dm"], rho, u, fi, f)[:44]
                elif what[0] == "tail":
                    data = tail
                elif what[0] == "size":
                    data = np.array([Nx * Ny * 8 * (3 if what[1] == 2 else 1)], "<u8").tobytes()
                elif what[0] == "records":
                    data = rec.tobytes()
                else:
                    _, blk, j = what
                    w = (e - b) * (24 if blk == 2 else 8)
                    data = blocks[blk][j * w:(j + 1) * w]
                assert len(data) == n, (what, len(data), n)
                os.pwrite(fd, data, off)
            os.close(fd)
        dist.barrier()
        if rank == 0:
            want = F.vti_bytes(rho, u, scal["Dx"], scal["Dt"], scal["Dm"], scal["Drho"], 1.0, 0.5)
            assert open(os.path.join(out_dir, "Fluid.3.vti"), "rb").read() == want
            want = F.restart_bytes(3, 1.7, scal["Dx"], scal["Dt"], scal["Dm"], rho, u, fi, f)
            assert open(os.path.join(out_dir, "Fluid.restart"), "rb").read() == want
            open(os.path.join(out_dir, "files_ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (12, 7)), (3, (13, 5))])
def test_ranks_write_one_file_together_over_gloo(world, shape, tmp_path, lib_built):
    """Every rank pwrite()s the byte ranges dist.file_plan gives it (what lbm_file.cu does with nranks > 1); the result must be
    the reference's file of the whole lattice, byte for byte."""
    import torch.multiprocessing as mp
    Nx, Ny = shape
    mp.spawn(_file_worker, args=(world, _free_port(), Nx, Ny, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "files_ok").exists()


def test_file_plan_tiles_the_file_exactly(lib_built):
    from life_b200 import capi, dist as D
    for Nx, Ny, world in ((12, 7, 2), (13, 5, 3), (64, 33, 8), (9, 4, 1)):
        head, tail = capi.vtk_frame(Nx, Ny, 0.5)
        for kind, total in (("vti", len(head) + 24 + 5 * 8 * Nx * Ny + len(tail)), ("restart", 44 + 120 * Nx * Ny)):
            spans = sorted((off, off + n) for r in range(world) for off, n, _ in D.file_plan(kind, Nx, Ny, world, r, len(head), len(tail)))
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0], (kind, Nx, Ny, world, a, b)      # no gap, no overlap


# ---- the in-place ("shift") layout of cfg.inplace, as a numpy model executed by real ranks -------------------------------------------
def _inplace_worker(rank, world, port, Nx, Ny, steps, out_dir):
    """One population buffer per rank: population v of logical element n sits at plane element (n - off_v) mod S (csrc/ctx.h:
    PopShift).  A step = every node keeps its nine values where they are (omega = 0: collision is the identity), the offsets grow by
    shift_v = cx*P + cy (that IS the push), then y-wrap, x-halo through contiguous staging (the planes are circular, halo.cu:
    k_halo_pack / k_halo_unpack) — all in the logical view.  After `steps` steps over gloo the lattice must equal `steps` applications
    of the reference's push map (src/Grid.cpp:229,240)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from life_b200 import capi, dist as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        b, e = capi.slab_range(Nx, world, rank)
        nxl = e - b
        JOFF = 2
        P = Ny + 4                      # column pitch with ghost rows at JOFF-1 and JOFF+Ny (the library pads to a multiple of 16)
        S = (nxl + 2) * P               # plane size, ghost columns 0 and nxl+1
        planes = np.zeros((9, S))
        off = np.zeros(9, dtype=np.int64)
        at = lambda v, idx: (idx - off[v]) % S                       # PopShift::at
        node = lambda c, r: c * P + r
        tags = np.arange(1, Nx * Ny * 9 + 1, dtype=np.float64).reshape(Nx, Ny, 9)
        cols = np.arange(1, nxl + 1)[:, None]
        rows = (JOFF + np.arange(Ny))[None, :]
        own = node(cols, rows)                                        # logical element of every owned node
        for v in range(9):
            planes[v, at(v, own)] = tags[b:e, :, v]                  # upload: plain layout (offsets zero)
        plan = D.halo_plan(Nx, Ny, world, periodic_x=True)
        allc = np.arange(nxl + 2)
        for _ in range(steps):
            # the sweep: in place, nothing moves; streaming = the offsets advance
            for v in range(9):
                off[v] = (off[v] + CX[v] * P + CY[v]) % S
            # y wrap of every column incl. the ghost columns (k_wrap_y), logical view
            for v in range(9):
                if CY[v] == 1:
                    planes[v, at(v, node(allc, JOFF))] = planes[v, at(v, node(allc, JOFF + Ny))]
                elif CY[v] == -1:
                    planes[v, at(v, node(allc, JOFF + Ny - 1))] = planes[v, at(v, node(allc, JOFF - 1))]
            # x halo: gather the ghost columns into contiguous buffers, ship, scatter into the edge columns
            r = JOFF + np.arange(Ny)
            reqs, bufs = [], []
            for (src, dst, pops, n) in plan:
                if src == rank:
                    col = nxl + 1 if pops == (1, 5, 7) else 0
                    msg = torch.from_numpy(np.stack([planes[v, at(v, node(col, r))] for v in pops]).copy())
                    reqs.append(dist.isend(msg, dst=dst, tag=0 if pops == (1, 5, 7) else 1))
                if dst == rank:
                    buf = torch.empty((3, Ny), dtype=torch.float64)
                    reqs.append(dist.irecv(buf, src=src, tag=0 if pops == (1, 5, 7) else 1))
                    bufs.append((pops, buf))
            for q in reqs:
                q.wait()
            for pops, buf in bufs:
                col = 1 if pops == (1, 5, 7) else nxl
                for k, v in enumerate(pops):
                    planes[v, at(v, node(col, r))] = buf[k].numpy()
        got = np.stack([planes[v, at(v, own)] for v in range(9)], axis=-1)       # download through the shifted layout (k_pack)
        full = D.gather_slabs(np.ascontiguousarray(got), Nx)
        if rank == 0:
            expect = tags
            i, j = np.meshgrid(np.arange(Nx), np.arange(Ny), indexing="ij")
            for _ in range(steps):
                nxt = np.zeros_like(expect)
                for v in range(9):
                    nxt[(i + CX[v] + Nx) % Nx, (j + CY[v] + Ny) % Ny, v] = expect[i, j, v]
                expect = nxt
            assert np.array_equal(full, expect)
            open(os.path.join(out_dir, "ok_inplace"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,steps", [(2, (12, 7), 5), (3, (13, 5), 4), (2, (9, 16), 7)])
def test_inplace_layout_design_over_gloo(world, shape, steps, tmp_path, lib_built):
    import torch.multiprocessing as mp
    Nx, Ny = shape
    mp.spawn(_inplace_worker, args=(world, _free_port(), Nx, Ny, steps, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok_inplace").exists()
