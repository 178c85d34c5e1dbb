"""bench.py --gpus N (N > 1): the slab-decomposed right-hand side on the GPUs of one box.

Launched by torchrun (one rank per GPU, NCCL for set-up and barriers, peer memory for the ghosts).
Headline = BASELINE config 4, weak scaling: the 3-D dam-break geometry at a particle spacing that
gives about `--per-gpu` (default 12.5 M) fluid particles per GPU -- 100 M on 8 -- cut along x into N
slabs of equal fluid count.  Every rank generates only its own slab of the lattice
(`slabs.dam_break_3d_slab`).  Before anything is timed, every run checks the slab path against a
single-handle kick of a small perturbed lattice on the same GPUs (`parity_check`); a failure exits
non-zero.  `variants`: the same at about 1 M fluid particles per GPU (the round-1 headline) and
strong scaling of the 10 M lattice over the N GPUs."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np


def _bind_to_gpu_numa_node(index: int) -> str:
    """Run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return f"{len(os.sched_getaffinity(0))} cpus (nvmlDeviceSetCpuAffinity)"
    except Exception as err:  # not fatal: the numbers are then what an unpinned process gets
        return f"unpinned ({type(err).__name__})"


def dx_for(n_fluid_total: float) -> float:
    """Particle spacing of the dam-break column (2 x 1 x 1) holding about n_fluid_total particles."""
    return float((2.0 / n_fluid_total) ** (1.0 / 3.0))


# ------------------------------------------------------------------ parity before timing
def parity_check(rank, world, local_rank, variant=0):
    """Small perturbed 3-D dam break (global lattice on every rank): the rank's slab dv over peer
    memory must equal its slab dv over NCCL send/recv bit for bit, and both must agree with the rows
    of a single-handle kick of the whole lattice on this rank's own GPU to 1e-5 (Float32; the slab
    sums neighbours in another order).  Returns the record for the JSON line."""
    import torch
    import torch.distributed as dist
    import trixiparticles.jl_b200 as tp
    from trixiparticles.jl_b200 import examples
    from trixiparticles.jl_b200.slabs import SlabSemidiscretization

    dev = torch.device("cuda", local_rank)
    dx = 0.025 if world > 4 else 0.04      # 80 / 50 fluid columns: every slab wider than the ghost layer
    fluid, wall, _ = examples.dam_break_3d(dx)
    u, v = examples.perturbed_state(fluid)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(
        device=local_rank, ode_memory="device", interact_variant=variant))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    u_d, v_d = torch.from_numpy(u.reshape(-1)).to(dev), torch.from_numpy(v.reshape(-1)).to(dev)
    ref = torch.full_like(v_d, float("nan"))
    ode.f1(ref, v_d, u_d, ode.p, 0.0)
    semi.synchronize()
    ref = ref.cpu().numpy().reshape(v.shape)
    semi.close()
    got, transports = {}, {}
    saved = os.environ.get("TPB_HALO")
    for mode in ("peer", "nccl"):
        os.environ["TPB_HALO"] = mode
        slab = SlabSemidiscretization(fluid, wall, rank=rank, world=world, device=local_rank,
                                      interact_variant=variant)
        so = slab.semidiscretize((0.0, 1.0))
        so.u0.copy_(torch.from_numpy(u[slab.owned_index].reshape(-1)).to(dev))
        so.v0.copy_(torch.from_numpy(v[slab.owned_index].reshape(-1)).to(dev))
        dv = torch.full_like(so.v0, float("nan"))
        for _ in range(3):      # both parities of the receive areas
            so.f1(dv, so.v0, so.u0, so.p, 0.0)
        slab.semi.synchronize()
        got[mode] = dv.cpu().numpy().reshape(-1, v.shape[1])
        transports[mode] = slab.halo_transport
        owned = slab.owned_index.copy()
        n_ghost = slab.n_ghost
        dist.barrier()
        slab.close()
    if saved is None:
        os.environ.pop("TPB_HALO", None)
    else:
        os.environ["TPB_HALO"] = saved
    r = ref[owned]
    err_acc = float(np.abs(got["peer"][:, :3] - r[:, :3]).max() / np.abs(ref[:, :3]).max())
    err_rho = float(np.abs(got["peer"][:, 3] - r[:, 3]).max() / np.abs(ref[:, 3]).max())
    identical = bool(np.array_equal(got["peer"], got["nccl"]))
    finite = bool(np.isfinite(got["peer"]).all())
    stats = torch.tensor([err_acc, err_rho, 0.0 if identical else 1.0, 0.0 if finite else 1.0,
                          0.0 if transports["peer"].startswith("peer") else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([float(len(owned)), float(n_ghost)], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    tol = 1e-5
    ok = bool(stats[0] <= tol and stats[1] <= tol and stats[2] == 0 and stats[3] == 0 and
              int(cnt[0]) == fluid.nparticles)
    return {"ok": ok, "workload": f"dam_break_3d dx = {dx} perturbed (SURVEY 8(d) M5), {fluid.nparticles} fluid + "
                                  f"{wall.nparticles} wall, {world} slabs",
            "max_rel_err_acceleration_vs_single_gpu": float(stats[0]),
            "max_rel_err_drho_vs_single_gpu": float(stats[1]), "tolerance": tol,
            "peer_memory_equals_nccl_bitwise": bool(stats[2] == 0),
            "peer_memory_transport_used": bool(stats[4] == 0),
            "owned_rows_total": int(cnt[0]), "ghost_rows_total": int(cnt[1])}


# ------------------------------------------------------------------ one timed configuration
def timed_slab_run(B, args, rank, world, local_rank, dx, steps, warm, e2e=True, clocks=None):
    import torch
    import torch.distributed as dist
    from trixiparticles.jl_b200.slabs import SlabSemidiscretization, dam_break_3d_slab

    dev = torch.device("cuda", local_rank)
    t_setup = time.perf_counter()
    fluid_k, wall_k, local = dam_break_3d_slab(dx, rank, world)
    n_f = local["n_fluid"]
    slab = SlabSemidiscretization(fluid_k, wall_k, rank=rank, world=world, device=local_rank,
                                  interact_variant=args.variant, local=local)
    n_w_local = wall_k.nparticles
    del fluid_k, wall_k
    ode = slab.semidiscretize((0.0, 1.0))
    t_setup = time.perf_counter() - t_setup
    u_d, v_d = ode.u0, ode.v0
    dv_d, du_d = torch.empty_like(v_d), torch.empty_like(u_d)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    if clocks is not None and rank == 0:
        clocks.start()
    for _ in range(warm):
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
    slab.semi.synchronize()
    st0 = slab.semi.stats()
    halo_launches0 = slab.peer.launches if slab.peer is not None else 0
    slab.semi.set_profiling(steps)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        flush.fill_(k & 0xFF)
        starts[k].record()
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)     # halo exchange + rebuild + Adami + interact
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        ends[k].record()
    torch.cuda.synchronize()
    dist.barrier()
    t_bracket = time.perf_counter() - t0
    slab.semi.synchronize()
    ms_steps = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    phases = slab.semi.phase_times()
    st1 = slab.semi.stats()
    launches = int(st1.kernel_launches_total - st0.kernel_launches_total)
    if slab.peer is not None:
        launches += slab.peer.launches - halo_launches0
    assert bool(torch.isfinite(dv_d).all()), "non-finite dv in the timed region"

    tt = torch.tensor([float(ms_steps.sum()), phases["interact"], float(slab.n_owned), float(slab.n_ghost),
                       float(n_w_local), phases["rebuild"], phases["boundary"], float(np.median(ms_steps)),
                       float(ms_steps.min())], dtype=torch.float64, device=dev)
    tmax, tmin = tt.clone(), tt.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    total_ms = float(tmax[0])
    clk = None
    if clocks is not None:
        # keep the same load running long enough for the 200 ms clock samples (the same number of
        # steps on every rank: the kicks exchange ghosts)
        tail = max(0, int(np.ceil(2000.0 / max(total_ms / steps, 1e-3))) - steps - warm)
        for _ in range(min(tail, 4000)):
            ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
            ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        torch.cuda.synchronize()
        clk = clocks.stop() if rank == 0 else None
    out = {"dx": dx, "n_fluid": n_f, "value": n_f * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps,
           "ms_per_step_median_slowest_rank": float(tmax[7]), "ms_per_step_min_fastest_rank": float(tmin[8]),
           "steps": steps, "n_fluid_per_rank": [int(tmin[2]), int(tmax[2])],
           "n_ghost_per_rank": [int(tmin[3]), int(tmax[3])], "n_wall_per_rank": [int(tmin[4]), int(tmax[4])],
           "interact_ms_slowest_rank": float(tmax[1]), "rebuild_ms_slowest_rank": float(tmax[5]),
           "boundary_ms_slowest_rank": float(tmax[6]), "bracket_s": t_bracket, "setup_s": t_setup,
           "launches_per_rank": launches, "halo": slab.halo_transport, "clocks": clk,
           "phases_ms_rank0": {k: phases[k] for k in ("rebuild", "density", "boundary", "interact")}}

    if e2e:
        # e2e: the rank's ODE vectors live in pinned host memory next to its GPU; per step u and v
        # are copied in, dv and du are written by the kernels straight into the (mapped) host
        # buffers, so the device->host transfer rides along with the computation
        n0, nd, nv = slab.n_owned, slab.nd, slab.nv
        hu = torch.empty(u_d.shape, dtype=u_d.dtype, pin_memory=True).copy_(u_d)
        hv = torch.empty(v_d.shape, dtype=v_d.dtype, pin_memory=True).copy_(v_d)
        hdv = torch.empty(v_d.shape, dtype=v_d.dtype, pin_memory=True)
        hdu = torch.empty(u_d.shape, dtype=u_d.dtype, pin_memory=True)

        def e2e_step():
            u_d.copy_(hu, non_blocking=True)
            v_d.copy_(hv, non_blocking=True)
            ode.f1(hdv, v_d, u_d, ode.p, 0.0)      # unified addressing: the pinned buffer is the dv pointer
            ode.f2(hdu, v_d, u_d, ode.p, 0.0)
            torch.cuda.synchronize()

        for _ in range(3):
            e2e_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        dist.barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        assert bool(torch.isfinite(hdv).all()) and torch.equal(hdu.view(n0, nd), hv.view(n0, nv)[:, :nd].to(hdu.dtype))
        nbytes = torch.tensor([float(hu.numel() * hu.element_size() + hv.numel() * hv.element_size()),
                               float(hdv.numel() * hdv.element_size() + hdu.numel() * hdu.element_size())],
                              dtype=torch.float64, device=dev)
        dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
        out["e2e"] = {"value": n_f * steps / float(te[0]), "unit": B.UNIT, "h2d_bytes_per_step": int(nbytes[0]),
                      "d2h_bytes_per_step": int(nbytes[1]), "ms_per_step": 1e3 * float(te[0]) / steps,
                      "timing": "host wall clock between barriers, max over ranks; pinned host ODE vectors per rank "
                                "(H2D copies of u, v; dv, du stored by the kernels into mapped host memory)"}
        del hu, hv, hdv, hdu
    slab.close()
    del u_d, v_d, dv_d, du_d, flush, ode
    torch.cuda.empty_cache()
    return out


def single_gpu_reference(B, args, local_rank, dx, steps, warm):
    """The same lattice family on ONE GPU (rank 0 only, the other ranks wait): the denominator of the
    weak-scaling efficiency at the per-GPU size of the headline run."""
    import torch
    import trixiparticles.jl_b200 as tp
    from trixiparticles.jl_b200 import examples
    dev = torch.device("cuda", local_rank)
    fluid, wall, _ = examples.dam_break_3d(dx)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(
        device=local_rank, ode_memory="device", interact_variant=args.variant))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    n_f, n_w = fluid.nparticles, wall.nparticles
    del fluid, wall
    u_d, v_d = ode.u0, ode.v0
    dv_d, du_d = torch.empty_like(v_d), torch.empty_like(u_d)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
    semi.set_profiling(steps)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        starts[k].record()
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        ends[k].record()
    torch.cuda.synchronize()
    semi.synchronize()
    ms = float(np.mean([s.elapsed_time(e) for s, e in zip(starts, ends)]))
    phases = semi.phase_times()
    semi.close()
    return {"dx": dx, "n_fluid": n_f, "n_wall": n_w, "ms_per_step": ms, "value": n_f / (ms * 1e-3), "steps": steps,
            "phases_ms": {k: phases[k] for k in ("rebuild", "density", "boundary", "interact")}}


def run(args):
    import torch
    import torch.distributed as dist
    import bench as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N > 1 must be launched with torchrun "
                         "(python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...)")
    affinity = _bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    ex, _ = B.WORKLOADS[args.workload]
    if ex != "dam_break_3d":
        raise SystemExit("the slab-decomposed bench runs the 3-D dam break family")

    # ---- parity first: nothing is timed on a path that does not reproduce the single-GPU kick
    parity = parity_check(rank, world, local_rank, args.variant)
    if not parity["ok"]:
        if rank == 0:
            print(json.dumps({"parity_check": parity, "error": "slab path does not reproduce the single-GPU kick"}))
        dist.destroy_process_group()
        sys.exit(3)

    warm = max(args.warmup, 3)
    per_gpu = float(args.per_gpu)
    dx_main = dx_for(per_gpu * world)
    clocks = B.ClockSampler(local_rank)
    main = timed_slab_run(B, args, rank, world, local_rank, dx_main, args.steps, warm, e2e=True, clocks=clocks)
    variants = {}
    if not args.no_variants:
        vsteps = min(args.steps, 20)
        # the round-1 headline: about 1 M fluid particles per GPU
        variants["weak_1m_per_gpu"] = timed_slab_run(B, args, rank, world, local_rank,
                                                     B.WORKLOADS["dam_break_3d_1m"][1] / world ** (1.0 / 3.0),
                                                     vsteps, warm, e2e=True)
        # strong scaling: the 10 M lattice of the single-GPU run over N GPUs
        variants["strong_10m_total"] = timed_slab_run(B, args, rank, world, local_rank,
                                                      B.WORKLOADS["dam_break_3d_10m"][1], vsteps, warm, e2e=False)
        # one GPU at the per-GPU size of the headline (weak-scaling denominator), rank 0 only
        if rank == 0:
            variants["single_gpu_same_size_per_gpu"] = single_gpu_reference(B, args, local_rank, dx_for(per_gpu),
                                                                            vsteps, warm)
            variants["single_gpu_10m"] = single_gpu_reference(B, args, local_rank,
                                                              B.WORKLOADS["dam_break_3d_10m"][1], vsteps, warm)
        dist.barrier()

    if rank == 0:
        peak_gbs, peak_src, sm_max_mhz = B.load_peaks()
        bpp = B.bytes_per_particle(3, 4, 4)
        k_ms = main["interact_ms_slowest_rank"]
        k_bytes = float(main["n_fluid_per_rank"][1]) * bpp["kernel_fluid"]
        achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        step_bytes = float(main["n_fluid"]) * bpp["step_fluid"]
        step_achieved = step_bytes / (main["ms_per_step"] * 1e-3) / 1e9
        eff = None
        if "single_gpu_same_size_per_gpu" in variants:
            eff = main["value"] / (world * variants["single_gpu_same_size_per_gpu"]["value"])
            s10 = variants["single_gpu_10m"]["value"]
            variants["strong_10m_total"]["efficiency_vs_single_gpu_10m"] = variants["strong_10m_total"]["value"] / (world * s10)
            e1 = variants["weak_1m_per_gpu"]
            variants["weak_1m_per_gpu"]["e2e_over_device"] = e1["e2e"]["value"] / e1["value"]
        line = {
            "metric": B.METRIC, "value": main["value"], "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"dam_break_3d geometry at dx = {dx_main:.6g}: {main['n_fluid']} fluid particles "
                                   f"in {world} slabs along x (BASELINE config 4, {per_gpu / 1e6:g} M per GPU)",
                       "n_fluid": main["n_fluid"], "n_fluid_per_rank": main["n_fluid_per_rank"],
                       "n_ghost_per_rank": main["n_ghost_per_rank"], "n_wall_per_rank": main["n_wall_per_rank"],
                       "ndims": 3, "nhs": "rebuilt every kick",
                       "halo": f"R + skin, 2R + skin along the walls; every kick over {main['halo']}",
                       "lattice": "each rank generates its own slab (RectangularTank x_window)",
                       "l2": "flushed between steps (256 MiB write, untimed)",
                       "timing": "sum of per-step CUDA-event times, max over ranks",
                       "bracket_s": main["bracket_s"], "setup_s": main["setup_s"], "cpu_affinity": affinity},
            "parity_check": parity,
            "clocks": main["clocks"], "e2e": main["e2e"],
            "gpu_launches": main["launches_per_rank"] * world,
            "roofline": {"bound": "hbm", "bound_measured": "smem_pipe",
                         "bound_actual": "shared-memory pipe / issue slots (pair sweep)",
                         "kernel": "interact! phase, slowest rank", "achieved": achieved,
                         "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "peak_source": peak_src,
                         "traffic": None, "kernel_ms": k_ms,
                         "step_achieved_all_ranks": step_achieved, "step_frac_per_gpu": step_achieved / world / peak_gbs,
                         "note": "FP32-issue / shared-memory bound pair sweep, see DESIGN.md section 3"},
            "phases_ms_slowest_rank": {"rebuild": main["rebuild_ms_slowest_rank"],
                                       "boundary": main["boundary_ms_slowest_rank"],
                                       "interact": main["interact_ms_slowest_rank"]},
            "step_ms": {"median_slowest_rank": main["ms_per_step_median_slowest_rank"],
                        "min_fastest_rank": main["ms_per_step_min_fastest_rank"]},
            "weak_scaling_efficiency_vs_single_gpu_same_size": eff,
            "e2e_over_device": main["e2e"]["value"] / main["value"],
            "variants": variants or None,
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
