"""Multi-GPU parity worker (run under torchrun by tests/test_gpu_multi.py, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/mp_parity.py CASE [STEPS]

Every rank builds the oracle's whole-lattice state (deterministic), takes its x-slab, steps it through the C ABI with the
NCCL halo exchange, and compares its slab with the oracle's whole-lattice result: fields and marker forces within relative
L2 1e-10.  For cases with bodies the recorded call-site trace of the compiled reference (tests/golden) drives the markers.
With a fourth argument (a directory visible to every rank) the ranks also write ONE Fluid.<t>.vti and ONE Fluid.restart
together (life_write_vtk / life_write_restart: every rank its own byte ranges), rank 0 checks the bytes against the reference
format built from the gathered slabs, and every rank reads its slab back with life_read_restart.
Exit code 0 = parity on every rank.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from life_b200 import capi, dist as D
    from tests import cases as K

    case = sys.argv[1]
    rank, world, local = D.env_ranks()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nid = D.share_nccl_id()

    g = K.golden(case)
    steps = (int(sys.argv[2]) if len(sys.argv) > 2 else 0) or int(g["steps"])
    o = K.make_oracle(g)
    extra = dict(inplace=1) if os.environ.get("LIFE_TEST_INPLACE") == "1" else {}     # cfg.inplace: one population buffer, shifted layout
    cfg = K.life_config(o.params, o, rank=rank, nranks=world, device=local, **extra)
    ctx = capi.Context(cfg, nccl_id=nid)
    Nx = o.Nx
    b, e = ctx.i_begin, ctx.i_end
    assert (b, e) == capi.slab_range(Nx, world, rank)
    sl = lambda name: D.slab_of(o.get(name), Nx, rank, world)
    ctx.upload_state(sl("f"), sl("rho"), sl("u"), sl("force_xy"), sl("force_ibm"), o.get("u_in"), o.get("rho_in"))

    has_ibm = "trace_step" in g.files
    worst_force = 0.0
    if has_ibm:
        k = 0
        for t in range(1, steps + 1):
            ctx.step(t)
            o.t = t
            o.lbm_kernel()
            while True:
                assert g["trace_step"][k] == t
                ctx.ibm_set_markers(g["trace_pos"][k], g["trace_vel"][k], g["trace_ds"][k], g["trace_eps"][k])
                force = ctx.ibm_interp()
                worst_force = max(worst_force, K.rel_l2(force, g["trace_force"][k], floor=1e-6))
                # the oracle is driven by the same recorded host state, so whole slabs can be compared below
                o.set_markers(g["trace_pos"][k], g["trace_vel"][k], g["trace_ds"][k], g["trace_eps"][k])
                o.find_support()
                o.ibm_interp()
                last = g["trace_last"][k]
                k += 1
                if last:
                    break
            ctx.ibm_spread()
            o.ibm_spread()
    else:
        for t in range(1, steps + 1):
            ctx.step(t)
        o.step(steps)
    vmax, has_nan, _, _ = ctx.max_speed()
    st = ctx.download_state()

    ok = True
    msgs = []
    if len(sys.argv) > 3:
        from tests import restartfile as R, vtkfile as V
        out = sys.argv[3]
        vti, rst = os.path.join(out, "Fluid.%d.vti" % steps), os.path.join(out, "Fluid.restart")
        slabs = [None] * world
        dist.all_gather_object(slabs, st)
        for mode in (capi.IO_SYNC, capi.IO_ASYNC):
            ctx.write_vtk(vti, o.params.rho_p, 0.5, mode)
            ctx.write_restart(rst, steps, mode)
            ctx.io_wait()                      # collective: all bytes are in the files, rank 0 has renamed the restart file
            if rank == 0:
                whole = {k: np.concatenate([s_[k] for s_ in slabs], axis=0) for k in st}
                want_vti = V.fluid_bytes(whole["rho"], whole["u"], o.Dx, o.Dt, o.Dm, o.Drho, o.params.rho_p, 0.5)
                want_rst = R.fluid_bytes(steps, o.params.omega, o.Dx, o.Dt, o.Dm, whole["rho"], whole["u"], whole["force_ibm"], whole["f"])
                if open(vti, "rb").read() != want_vti:
                    ok = False
                    msgs.append("shared .vti differs (mode %d)" % mode)
                if open(rst, "rb").read() != want_rst or os.path.exists(rst + ".temp"):
                    ok = False
                    msgs.append("shared Fluid.restart differs (mode %d)" % mode)
            dist.barrier()
        back = capi.Context(K.life_config(o.params, o, rank=rank, nranks=world, device=local, **extra), nccl_id=D.share_nccl_id())
        t_file = back.read_restart(rst, o.get("force_xy").reshape(-1, 2)[0], o.get("u_in"), o.get("rho_in"))
        sb = back.download_state()
        back.close()
        if t_file != steps or any(not np.array_equal(sb[k], st[k]) for k in st):
            ok = False
            msgs.append("life_read_restart did not return this rank's slab")
    ctx.close()
    if has_ibm:
        if not worst_force < K.TOL:
            ok = False
            msgs.append("marker force %.3e" % worst_force)
        # golden fixture: sampled nodes of the compiled reference's final state; compare those that fall in this slab
        s = g["sample"]
        i, j = s // int(g["Ny"]), s % int(g["Ny"])
        m = (i >= b) & (i < e)
        for name in ("rho", "u", "f", "force_ibm"):
            if m.any():
                err = K.rel_l2(st[name][i[m] - b, j[m]], g[name][m], floor=1e-12 if name == "force_ibm" else 0.0)
                if not err < K.TOL:
                    ok = False
                    msgs.append("%s %.3e" % (name, err))
        for name in ("rho", "u", "f", "force_ibm"):
            err = K.rel_l2(st[name], o.get(name)[b:e], floor=1e-12 if name == "force_ibm" else (1e-5 if name == "u" else 0.0))
            if not err < K.TOL:
                ok = False
                msgs.append("slab %s %.3e" % (name, err))
    else:
        for name in ("rho", "u", "f"):
            # a slab can lie where the flow has not arrived yet (ChannelFlow downstream of the inlet front): u = sum(c f)/rho
            # there is the cancellation noise of O(0.1) populations, ~1e-16 absolute in the reference too, so the
            # denominator is floored at 1e-5 per entry (error bar 1e-15 absolute)
            err = K.rel_l2(st[name], o.get(name)[b:e], floor=1e-5 if name == "u" else 0.0)
            if not err < K.TOL:
                ok = False
                msgs.append("%s %.3e" % (name, err))
        u = o.get("u")
        want = np.sqrt((u ** 2).sum(axis=-1)).max()
        if has_nan or abs(vmax - want) > 1e-12:
            ok = False
            msgs.append("max speed %r vs %r" % (vmax, want))
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    print("[rank %d/%d] %s columns [%d,%d): %s" % (rank, world, case, b, e, "ok" if ok else "MISMATCH " + "; ".join(msgs)), flush=True)
    dist.destroy_process_group()
    return int(flag.item() != 0)


if __name__ == "__main__":
    sys.exit(main())
