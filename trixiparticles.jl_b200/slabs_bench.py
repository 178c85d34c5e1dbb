"""bench.py --gpus N (N > 1): weak-scaling run of the slab-decomposed right-hand side.

Launched by torchrun (one rank per GPU, NCCL).  Every rank gets a slab holding about as many
fluid particles as the whole N = 1 workload: the same dam-break geometry at a particle spacing
of dx_1 / N^(1/3), cut along x into N slabs of equal fluid count."""
from __future__ import annotations

import json
import os
import time

import numpy as np


def run(args):
    import torch
    import torch.distributed as dist
    import bench as B
    import trixiparticles.jl_b200 as tp
    from trixiparticles.jl_b200 import examples
    from trixiparticles.jl_b200.slabs import SlabSemidiscretization

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N > 1 must be launched with torchrun "
                         "(python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)

    ex, dx1 = B.WORKLOADS[args.workload]
    if ex != "dam_break_3d":
        raise SystemExit("the slab-decomposed bench runs the 3-D dam break family")
    dx = dx1 / world ** (1.0 / 3.0)
    fluid, wall, _ = examples.dam_break_3d(dx)
    n_f, n_w = fluid.nparticles, wall.nparticles
    slab = SlabSemidiscretization(fluid, wall, rank=rank, world=world, device=local_rank,
                                  interact_variant=args.variant)
    del fluid, wall
    ode = slab.semidiscretize((0.0, 1.0))
    u_d, v_d = ode.u0, ode.v0
    dv_d, du_d = torch.empty_like(v_d), torch.empty_like(u_d)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    warm = max(args.warmup, 3)
    # the timed region is a few tens of milliseconds: sample the clocks from the warm-up on
    clocks = B.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(warm):
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
    slab.semi.synchronize()
    st0 = slab.semi.stats()
    slab.semi.set_profiling(args.steps)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
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
    # the timed region is too short for a 200 ms clock sample: keep the same load running for a
    # fixed number of steps (the same on every rank -- the kicks exchange ghosts) before reading it
    if args.steps < 800:
        for _ in range(800):
            ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
            ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        torch.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None

    # max over ranks of the device time of the K steps
    tt = torch.tensor([float(ms_steps.sum()), phases["interact"], float(slab.n_owned), float(slab.n_ghost),
                       float(slab.wall.nparticles if slab.wall is not None else 0)], dtype=torch.float64, device=dev)
    tmax = tt.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = tt.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    total_ms = float(tmax[0])
    value = n_f * args.steps / (total_ms * 1e-3)

    # e2e: owned ODE vectors in pinned host memory, copied in and out every step
    hu = torch.empty(u_d.shape, dtype=u_d.dtype, pin_memory=True).copy_(u_d)
    hv = torch.empty(v_d.shape, dtype=v_d.dtype, pin_memory=True).copy_(v_d)
    hdv = torch.empty(v_d.shape, dtype=v_d.dtype, pin_memory=True)
    hdu = torch.empty(u_d.shape, dtype=u_d.dtype, pin_memory=True)

    def e2e_step():
        u_d.copy_(hu, non_blocking=True)
        v_d.copy_(hv, non_blocking=True)
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        hdv.copy_(dv_d, non_blocking=True)
        hdu.copy_(du_d, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dist.barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(hu.numel() * hu.element_size() + hv.numel() * hv.element_size()),
                        float(hdv.numel() * hdv.element_size() + hdu.numel() * hdu.element_size())],
                       dtype=torch.float64, device=dev)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)

    launches = int(st1.kernel_launches_total - st0.kernel_launches_total)
    if rank == 0:
        peak_gbs, peak_src, _ = B.load_peaks()
        bpp = B.bytes_per_particle(3, 4, 4)
        k_ms = float(tmax[1])
        k_bytes = float(tmax[2]) * bpp["kernel_fluid"]
        achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} geometry at dx = {dx:.6g} ({world} slabs along x)",
                       "n_fluid": n_f, "n_wall": n_w, "n_fluid_per_rank": [int(tmin[2]), int(tmax[2])],
                       "n_ghost_per_rank": [int(tmin[3]), int(tmax[3])],
                       "n_wall_per_rank": [int(tmin[4]), int(tmax[4])], "ndims": 3,
                       "nhs": "rebuilt every kick", "halo": f"R + skin, 2R + skin along the walls; every kick over {slab.halo_transport}",
                       "l2": "flushed between steps (256 MiB write, untimed)",
                       "timing": "sum of per-step CUDA-event times, max over ranks",
                       "bracket_s": t_bracket},
            "clocks": clk,
            "e2e": {"value": n_f * args.steps / float(te[0]), "unit": B.UNIT,
                    "h2d_bytes_per_step": int(h2d[0]), "d2h_bytes_per_step": int(h2d[1]),
                    "ms_per_step": 1e3 * float(te[0]) / args.steps},
            "gpu_launches": launches * world,
            "roofline": {"bound": "hbm", "kernel": "interact! phase, slowest rank", "achieved": achieved,
                         "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "peak_source": peak_src,
                         "traffic": None, "kernel_ms": k_ms,
                         "note": "FP32-issue / shared-memory bound pair sweep, see DESIGN.md section 3"},
            "phases_ms_rank0": {k: phases[k] for k in ("rebuild", "density", "boundary", "interact")},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    slab.close()
    dist.destroy_process_group()
