#!/usr/bin/env python
"""bench.py -- particle-RHS evaluations per second of the WCSPH right-hand side on B200.

One "step" = one `kick!` + one `drift!` (TrixiParticles.jl semidiscretization.jl:522-612:
neighbourhood-search rebuild, EOS, Adami wall pressure, fluid-fluid and fluid-wall
`interact!`, gravity, du = v) over one synthetic particle lattice.  Metric (BASELINE.json):
particle-RHS evaluations / s = N_fluid * steps / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Own arm (default): the CUDA library through the reference-facing API.
  value     device-resident ODE vectors (torch CUDA tensors), CUDA events, L2 flushed
            between steps.
  e2e       the same step through `B200Backend(ode_memory="host")`: pinned host ODE vectors,
            host<->device copies inside the timed region.
  roofline  the dominant kernel phase (interact!) timed with CUDA events inside the library
            on the launching stream during the timed steps.
  cpu_baseline  the CPU oracle (a port of the reference algorithm; Julia is not installed)
            timed on this box's host cores on the same workload.
Reference arm (--impl reference): the CPU oracle port, all host threads, same workload.

N > 1 (torchrun): BASELINE config 4.  The lattice is slab-decomposed along x, one slab per rank
(generated locally by every rank), ghost particles are exchanged every kick over peer memory (one
pack-and-store kernel per neighbour; NCCL send/recv with TPB_HALO=nccl; see
trixiparticles.jl_b200/slabs.py and slabs_bench.py); "scaling" is "weak": every rank gets a slab of
`--per-gpu` (12.5 M) fluid particles -- 100 M on 8 GPUs.  The run starts with an in-process parity
check of the slab path against a single-GPU kick (`parity_check` in the JSON line; failure exits
non-zero) and carries the 1 M-per-GPU and the 10 M strong-scaling runs as `variants`.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-RHS evals/s (particles x steps/s)"
UNIT = "particle-RHS/s"

# name -> (example, particle spacing).  SURVEY.md section 8(d) M3 / M4.
WORKLOADS = {
    "dam_break_3d_1m": ("dam_break_3d", 0.0126),     # 992 319 fluid + 1 599 800 wall
    "dam_break_3d_10m": ("dam_break_3d", 0.00585),   # ~10 M fluid
    "dam_break_3d_12m5": ("dam_break_3d", 0.00543),  # ~12.5 M fluid: the per-GPU size of the --gpus N run
    "dam_break_3d_30m": ("dam_break_3d", 0.0040),    # SURVEY 8(d) M4 on ONE GPU: ~31 M fluid
    "dam_break_3d_100m": ("dam_break_3d", 0.0027),   # ~101 M fluid + 34 M wall on one GPU (records ~5 GB of the 180 GB)
    "dam_break_3d_250k": ("dam_break_3d", 0.02),     # reduced sample for slow CPU legs
    "dam_break_3d_small": ("dam_break_3d", 0.05),    # quick functional check
    "dam_break_2d": ("dam_break_2d", 40),            # config 1 (Float64)
}
DEFAULT_WORKLOAD = "dam_break_3d_1m"

# algorithmic bytes per particle (DESIGN.md "Roofline accounting")
def bytes_per_particle(nd, tsize, csize):
    pad16 = lambda b: (b + 15) // 16 * 16
    nv = nd + 1
    b_io = nd * csize + nv * tsize + tsize + nv * tsize + 2 * nd * tsize
    b_nhs = 8 + 2 * (pad16(nd * csize + tsize) + pad16(nv * tsize)) + 8
    step_fluid = b_io + b_nhs
    step_wall = 2 * pad16(nd * csize + tsize) + 2 * (2 * tsize)
    # interact! kernel alone: sorted records in (A, B, P, perm), dv out; wall tiles (A, W) in
    k_fluid = pad16(nd * csize + tsize) + pad16(nv * tsize) + tsize + 4 + nv * tsize
    k_wall = pad16(nd * csize + tsize) + 2 * tsize
    return dict(step_fluid=step_fluid, step_wall=step_wall, kernel_fluid=k_fluid, kernel_wall=k_wall)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def make_workload(name, eltype=None, coords=None, shuffle=False, adaptive=False, summation=False):
    from trixiparticles.jl_b200 import examples
    ex, arg = WORKLOADS[name]
    dt = {None: None, "f32": np.float32, "f64": np.float64}
    kw = {}
    if dt[eltype] is not None:
        kw["eltype"] = dt[eltype]
    if dt[coords] is not None:
        kw["coordinates_eltype"] = dt[coords]
    if ex == "dam_break_3d":
        if adaptive:
            kw["adaptive_sound_speed"] = True
        if summation:
            import trixiparticles.jl_b200 as tp_
            kw["density_calculator"] = tp_.SummationDensity()
        fluid, wall, _ = examples.dam_break_3d(arg, **kw)
    else:
        fluid, wall, _ = examples.dam_break_2d(arg, **kw)
    ic = fluid.initial_condition
    if shuffle:
        # random particle <-> ODE index map: what the reference's SortingCallback exists to undo
        # (callbacks/sorting.jl:17-19 quotes a 3-4x penalty for unsorted particles on its GPU path)
        perm = np.random.default_rng(1234).permutation(fluid.nparticles)
        for name_ in ("coordinates", "velocity", "mass", "density", "pressure"):
            setattr(ic, name_, np.ascontiguousarray(getattr(ic, name_)[perm]))
        fluid.mass = np.ascontiguousarray(fluid.mass[perm])
    u = np.ascontiguousarray(ic.coordinates, dtype=fluid.coordinates_eltype)
    cols = [ic.velocity] if summation else [ic.velocity, ic.density[:, None]]
    v = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=fluid.eltype)
    return fluid, wall, u, v


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if exe is None:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.out = open(self.path, "w")
        self.proc = subprocess.Popen([exe, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.device), "-lms", "200"], stdout=self.out,
                                     stderr=subprocess.DEVNULL)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ------------------------------------------------------------------ CPU legs (oracle port)
def cpu_step_time(fluid, wall, u, v, reps, nthreads=0):
    """Times `reps` oracle kick!+drift! evaluations; returns (seconds per step, threads)."""
    from oracle import adapter, oracle as O
    O.build()
    nd = fluid.ndims
    adapter.kick(fluid, wall, u, v, nthreads=nthreads, use_grid=2)  # warm-up (page faults, thread pool, static wall grid)
    t0 = time.perf_counter()
    for _ in range(reps):
        adapter.kick(fluid, wall, u, v, nthreads=nthreads, use_grid=2)
        O.drift(v, nd, u.dtype)
    dt = (time.perf_counter() - t0) / reps
    return dt, (nthreads or O.max_threads())


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia is not installed in
    this image, so the unmodified reference cannot run) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import adapter, oracle as O
    O.build()
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    _kick = adapter.kick
    adapter_kick = lambda *a, **k: _kick(*a, nthreads=nthreads, use_grid=2, **k)   # static wall grid built once
    name = args.workload
    fluid, wall, u, v = make_workload(name)
    nd = fluid.ndims
    t0 = time.perf_counter()
    adapter_kick(fluid, wall, u, v)
    t1 = time.perf_counter() - t0
    sample = f"full workload {name}: {fluid.nparticles} fluid + {wall.nparticles} wall particles per step"
    if (args.steps + args.warmup) * t1 > 240.0 and name != "dam_break_3d_250k":
        small = "dam_break_3d_250k"
        fluid, wall, u, v = make_workload(small)
        sample = (f"reduced lattice {small} of the same geometry ({fluid.nparticles} fluid + "
                  f"{wall.nparticles} wall per step) because the full {name} would exceed the time bound")
    for _ in range(args.warmup):
        adapter_kick(fluid, wall, u, v)
        O.drift(v, nd, u.dtype)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        adapter_kick(fluid, wall, u, v)
        O.drift(v, nd, u.dtype)
    dt = time.perf_counter() - t0
    value = fluid.nparticles * args.steps / dt
    cores = nthreads
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if np.dtype(fluid.eltype).itemsize == 4 else "f64", "data": "synthetic",
        "config": {"workload": name, "n_fluid": fluid.nparticles, "n_wall": wall.nparticles,
                   "note": "CPU oracle port of the reference algorithm (OpenMP); TrixiParticles.jl itself "
                           "needs Julia, which this image does not have"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ own arm, one GPU
def run_single(args):
    import torch
    import trixiparticles.jl_b200 as tp
    from trixiparticles.jl_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the accelerated path has no CPU fallback")
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    fluid, wall, u, v = make_workload(args.workload, args.eltype, args.coords, args.shuffle, summation=args.summation)
    if args.no_slip:
        # no-slip wall (`viscosity_wall = viscosity_fluid`) on a synthetic velocity field; not the headline
        wall.boundary_model.viscosity = fluid.viscosity
        v[:, :fluid.ndims] = (0.05 * np.sin(40.0 * u)).astype(v.dtype)
    nd, n_f, n_w = fluid.ndims, fluid.nparticles, wall.nparticles
    tsize, csize = np.dtype(fluid.eltype).itemsize, np.dtype(fluid.coordinates_eltype).itemsize
    if args.e2e_only:
        e2e = run_e2e(tp, torch, fluid, wall, u, v, args)
        e2e["env"] = {k: v for k, v in os.environ.items() if k.startswith("TPB_")}
        print(json.dumps(e2e))
        return

    # ---- device-resident arm
    backend = tp.B200Backend(device=0, ode_memory="device", interact_variant=args.variant)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=backend)
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    u_d = torch.from_numpy(u.reshape(-1)).to(dev)
    v_d = torch.from_numpy(v.reshape(-1)).to(dev)
    dv_d = torch.empty_like(v_d)
    du_d = torch.empty_like(u_d)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    if args.evolve > 0:
        # the state after `--evolve` steps of the library's own time loop
        # (CarpenterKennedy2N54, dt from StepsizeCallback(cfl = 0.9)): realistic disorder
        from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, StepsizeCallback, solve
        dt_ = StepsizeCallback(cfl=0.9).dt(semi)
        ode_e = tp.DynamicalODEProblem(ode.f1, ode.f2, v_d, u_d, (0.0, args.evolve * dt_), ode.p)
        sol = solve(ode_e, CarpenterKennedy2N54(williamson_condition=False), dt=dt_)
        u_d, v_d = sol.u.contiguous(), sol.v.contiguous()
        semi.synchronize()

    if args.sort:
        # SortingCallback (callbacks/sorting.jl): one tpb_sort_system undoes --shuffle; timed separately
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        semi.sort_particles(v_d, u_d)
        e1.record()
        torch.cuda.synchronize()
        sort_ms = e0.elapsed_time(e1)
    else:
        sort_ms = None

    def step():
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)

    for _ in range(max(args.warmup, 3)):
        step()
    semi.synchronize()
    torch.cuda.synchronize()
    st0 = semi.stats()
    semi.set_profiling(args.steps)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    clocks = ClockSampler(0)
    clocks.start()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)           # evict L2 (untimed: outside the event pair)
        starts[k].record()
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        mids[k].record()
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        ends[k].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    semi.synchronize()   # raises if a deferred device-side error (out of bounds) was flagged
    ms_steps = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    ms_kick = np.array([s.elapsed_time(m) for s, m in zip(starts, mids)])
    ms_drift = np.array([m.elapsed_time(e) for m, e in zip(mids, ends)])
    phases = semi.phase_times()
    st1 = semi.stats()
    # keep the GPU busy a little longer if the timed region was too short for a clock sample
    if t_wall < 1.0:
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize()
    clk = clocks.stop()
    total_ms = float(ms_steps.sum())
    value = n_f * args.steps / (total_ms * 1e-3)
    launches = int(st1.kernel_launches_total - st0.kernel_launches_total)

    if args.quick:
        print(json.dumps({"quick": True, "workload": args.workload, "ms_per_step": total_ms / args.steps,
                          "value": value, "phases_ms": phases, "kick_ms": float(ms_kick.mean()),
                          "variant": int(st1.interact_variant_used), "shuffle": bool(args.shuffle),
                          "sort_ms": sort_ms,
                          "env": {k: v for k, v in os.environ.items() if k.startswith("TPB_")}}))
        return

    # SURVEY 8(d) timing protocol: >= 200 timed kick!+drift! pairs, median and min (the driver picks
    # --steps; when it is smaller, the same loop simply runs on)
    proto = None
    n_proto = max(200, args.steps)
    if not args.no_variants:
        ps = [torch.cuda.Event(enable_timing=True) for _ in range(n_proto)]
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(n_proto)]
        for k in range(n_proto):
            flush.fill_(k & 0xFF)
            ps[k].record()
            ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
            ode.f2(du_d, v_d, u_d, ode.p, 0.0)
            pe[k].record()
        torch.cuda.synchronize()
        pm = np.array([a.elapsed_time(b) for a, b in zip(ps, pe)])
        proto = {"timed_evaluations": n_proto, "ms_median": float(np.median(pm)), "ms_min": float(pm.min()),
                 "ms_mean": float(pm.mean()), "value_median": n_f / (float(np.median(pm)) * 1e-3)}

    # L2-warm variant (no flush), reported for information
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_warm = e0.elapsed_time(e1) / args.steps

    # active wall particles = those with a fluid neighbour (Adami volume > 0)
    n_w_active = int((semi.system_field(wall, "volume") > 0).sum())
    # accepted pair counts (capacity 0 => only the count comes back)
    pairs = {}
    for name, a, b in (("fluid_fluid", fluid, fluid), ("fluid_wall", fluid, wall), ("wall_fluid", wall, fluid)):
        pairs[name] = semi.count_neighbor_pairs(a, b, u_d)

    # ---- roofline of the dominant kernel phase (interact!)
    peak_gbs, peak_src, sm_max_mhz = load_peaks()
    bpp = bytes_per_particle(nd, tsize, csize)
    k_bytes = n_f * bpp["kernel_fluid"] + n_w_active * bpp["kernel_wall"]
    k_ms = phases["interact"]
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    step_bytes = n_f * bpp["step_fluid"] + n_w_active * bpp["step_wall"]
    step_achieved = step_bytes / (total_ms / args.steps * 1e-3) / 1e9
    # algorithmic flops: 80 / 55 / 12 per accepted pair, 9 per candidate
    cand_per_particle = 729 if nd == 3 else 144 if args.workload == "dam_break_2d" else 9 * 16
    flops = (pairs["fluid_fluid"] * 80 + pairs["fluid_wall"] * 55 + pairs["wall_fluid"] * 12
             + 9.0 * cand_per_particle * (2 * n_f + n_w_active))
    fp32_peak = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    traffic, traffic_src, limiters = None, None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
        with open(tpath) as f:
            rec = json.load(f).get(args.workload)
        if rec:
            traffic, traffic_src = rec["dram_bytes_read"] + rec["dram_bytes_write"], rec["source"]
            limiters = {k: rec[k] for k in ("shared_memory_pipe_pct_of_peak", "issue_slots_pct_of_peak",
                                            "fma_pipe_pct_of_peak", "alu_pipe_pct_of_peak", "l2_hit_rate_pct",
                                            "limiters_note") if k in rec}
    fp32_tflops = flops / (ms_kick.mean() * 1e-3) / 1e12
    roofline = {
        # the contract's HBM figures (achieved / peak / frac, algorithmic bytes) ...
        "bound": "hbm",
        # ... and what ncu measures as the limiter of this kernel: first-class keys, same capture as `traffic`
        "bound_measured": "smem_pipe",
        "smem_pipe_frac": (limiters or {}).get("shared_memory_pipe_pct_of_peak", 0.0) / 100.0 or None,
        "issue_slot_frac": (limiters or {}).get("issue_slots_pct_of_peak", 0.0) / 100.0 or None,
        "bound_actual": "shared-memory pipe / issue slots: the pair sweep does about 270 flop "
                        "per compulsory byte, HBM cannot bound it (see limiters, fp32_frac)",
        "fp32_frac": fp32_tflops / fp32_peak, "step_frac": step_achieved / peak_gbs,
        "rebuild_frac": (n_f * 80 / (phases["rebuild"] * 1e-3) / 1e9 / peak_gbs) if phases["rebuild"] > 0 else None,
        "kernel": "interact! phase (fluid-fluid + fluid-wall pair sweep)",
        "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
        "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": k_bytes, "kernel_ms": k_ms,
        "kernel_share_of_step": k_ms / (total_ms / args.steps),
        "note": "the pair sweep is bound by the shared-memory pipe and the issue slots (about 300 flop per "
                "compulsory byte), see limiters, fp32 and DESIGN.md",
        "limiters": limiters,
        "step": {"achieved": step_achieved, "frac": step_achieved / peak_gbs, "algorithmic_bytes": step_bytes,
                 "bytes_per_fluid_particle": bpp["step_fluid"], "bytes_per_active_wall_particle": bpp["step_wall"]},
        "fp32": {"algorithmic_tflops": flops / (ms_kick.mean() * 1e-3) / 1e12, "nominal_peak_tflops": fp32_peak,
                 "frac_of_nominal": flops / (ms_kick.mean() * 1e-3) / 1e12 / fp32_peak,
                 "pairs": pairs},
    }
    semi.close()
    del u_d, v_d, dv_d, du_d, flush
    torch.cuda.empty_cache()

    # ---- e2e arm: host ODE vectors through the same API, copies inside the timed region
    e2e = run_e2e(tp, torch, fluid, wall, u, v, args)

    # ---- the reference's default single-precision set-up: Float32 fields, Float64 coordinates
    # (docs/src/gpu.md "Single precision simulations"; examples/fluid/dam_break_3d.jl:36-39)
    variants = None
    if args.eltype is None and args.coords is None and tsize == 4 and csize == 4 and not args.no_variants:
        variants = {"f32_fields_f64_coordinates": run_variant(tp, torch, args, "f32", "f64")}
        if WORKLOADS[args.workload][0] == "dam_break_3d":
            # the script as shipped: Float64 coordinates + StateEquationAdaptiveCole (one max|v|
            # reduction per kick; the new speed of sound stays on the device, k_adaptive_consts)
            variants["as_shipped_f64_coordinates_adaptive_cole"] = run_variant(tp, torch, args, "f32", "f64",
                                                                              adaptive=True)
        # no-slip wall (`viscosity_wall = viscosity_fluid`, examples/fluid/dam_break_2d.jl:78-80): the
        # wall-velocity sweep and the wall's viscous term run as two extra per-particle kernels
        # (both on the same synthetic velocity field, since the lattice itself is at rest)
        variants["moving_fluid_free_slip_wall"] = run_variant(tp, torch, args, "f32", "f32", moving=True)
        variants["moving_fluid_no_slip_wall"] = run_variant(tp, torch, args, "f32", "f32", moving=True,
                                                            no_slip=True)
        # SummationDensity (density_calculators.jl:26-50): one more tile sweep (W only) before interact!
        variants["summation_density"] = run_variant(tp, torch, args, "f32", "f32", summation=True)
        if args.workload == DEFAULT_WORKLOAD:
            # BASELINE config 4 on one GPU: the 10 M lattice (north_star's size) and the per-GPU size of
            # the weak-scaling run of `--gpus N` (12.5 M: the denominator of its efficiency)
            variants["dam_break_3d_10m"] = run_variant(tp, torch, args, "f32", "f32", workload="dam_break_3d_10m")
            variants["dam_break_3d_12m5"] = run_variant(tp, torch, args, "f32", "f32", workload="dam_break_3d_12m5")

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        dt, cores = cpu_step_time(fluid, wall, u, v, reps=1)
        reps = int(min(8, max(2, 12.0 / max(dt, 1e-3))))
        dt, cores = cpu_step_time(fluid, wall, u, v, reps=reps)
        cpu = {"value": n_f / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{reps} kick!+drift! evaluations of the full workload ({n_f} fluid + {n_w} wall) "
                         f"with the OpenMP CPU oracle (static wall cell list built once and reused, fluid cell list "
                         f"rebuilt every kick), {1e3 * dt:.1f} ms each"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if tsize == 4 else "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "n_fluid": n_f, "n_wall": n_w, "n_wall_active": n_w_active,
                   "ndims": nd, "coords_dtype": "f32" if csize == 4 else "f64",
                   "kernel": type(fluid.smoothing_kernel).__name__,
                   "nhs": "rebuilt every kick", "l2": "flushed between steps (256 MiB write, untimed)",
                   "interact_variant": int(st1.interact_variant_used),
                   **({"wall": "no-slip (wall viscosity = fluid viscosity), synthetic velocity field"}
                      if args.no_slip else {})},
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "phases_ms": {**{k: phases[k] for k in _lib.PHASES}, "kick": float(ms_kick.mean()),
                      "drift": float(ms_drift.mean()), "step_min": float(ms_steps.min()),
                      "step_median": float(np.median(ms_steps)), "step_l2_warm": ms_warm},
        "protocol_200": proto,
        "cpu_baseline": cpu, "variants": variants,
    }
    print(json.dumps(line))


def run_variant(tp, torch, args, eltype, coords, steps=10, adaptive=False, no_slip=False, moving=False,
                workload=None, summation=False):
    """Device-resident kick!+drift! of the same workload in another precision set-up (or of
    another lattice size: `workload`)."""
    fluid, wall, u, v = make_workload(workload or args.workload, eltype, coords, adaptive=adaptive,
                                      summation=summation)
    if no_slip:
        wall.boundary_model.viscosity = fluid.viscosity
    if moving:
        # the lattice is at rest; a velocity field exercises the viscous terms
        v = v.copy()
        v[:, :fluid.ndims] = (0.05 * np.sin(40.0 * u)).astype(v.dtype)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(
        device=0, ode_memory="device", interact_variant=args.variant))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dev = torch.device("cuda", 0)
    u_d, v_d = torch.from_numpy(u.reshape(-1)).to(dev), torch.from_numpy(v.reshape(-1)).to(dev)
    dv_d, du_d = torch.empty_like(v_d), torch.empty_like(u_d)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
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
    n_f = fluid.nparticles
    semi.close()
    out = {"ms_per_step": ms, "value": n_f / (ms * 1e-3), "unit": UNIT, "steps": steps,
           "dtype": eltype, "coords_dtype": coords, "phases_ms": {k: phases[k] for k in phases}}
    if workload is not None:
        peak_gbs, _, sm_max_mhz = load_peaks()
        bpp = bytes_per_particle(fluid.ndims, 4 if eltype == "f32" else 8, 4 if coords == "f32" else 8)
        out.update({"workload": workload, "n_fluid": n_f, "n_wall": wall.nparticles,
                    "roofline_step_frac_fluid_bytes_only": n_f * bpp["step_fluid"] / (ms * 1e-3) / 1e9 / peak_gbs,
                    "rebuild_gbs": n_f * 80 / (phases["rebuild"] * 1e-3) / 1e9 if phases.get("rebuild") else None})
    del u_d, v_d, dv_d, du_d, flush
    torch.cuda.empty_cache()
    return out


def run_e2e(tp, torch, fluid, wall, u, v, args):
    """Host-pointer mode of the C ABI (what a CPU-side integrator sees): pinned host ODE
    vectors, H2D of u and v and D2H of dv in `kick!`, H2D of v and D2H of du in `drift!`."""
    backend = tp.B200Backend(device=0, ode_memory="host", interact_variant=args.variant)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=backend)
    ode = tp.semidiscretize(semi, (0.0, 1.0))

    def pinned(a):
        t = torch.empty(a.size, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[:] = a.reshape(-1)
        return t

    tu, tv = pinned(u), pinned(v)
    tdv, tdu = pinned(np.zeros_like(v)), pinned(np.zeros_like(u))
    hu, hv, hdv, hdu = tu.numpy(), tv.numpy(), tdv.numpy(), tdu.numpy()
    for _ in range(max(args.warmup, 3)):
        ode.f1(hdv, hv, hu, ode.p, 0.0)
        ode.f2(hdu, hv, hu, ode.p, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ode.f1(hdv, hv, hu, ode.p, 0.0)   # returns after dv is in host memory
        ode.f2(hdu, hv, hu, ode.p, 0.0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert np.isfinite(hdv).all() and np.array_equal(hdu.reshape(u.shape), v[:, : fluid.ndims].astype(u.dtype))
    semi.close()
    return {"value": fluid.nparticles * args.steps / dt, "unit": UNIT,
            "h2d_bytes_per_step": int(hu.nbytes + 2 * hv.nbytes), "d2h_bytes_per_step": int(hdv.nbytes + hdu.nbytes),
            "ms_per_step": 1e3 * dt / args.steps, "timing": "host wall clock around the synchronous host-pointer calls"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=0, help="interact kernel variant (0 = auto)")
    ap.add_argument("--eltype", default=None, choices=["f32", "f64"], help="eltype(system) (default: the example's)")
    ap.add_argument("--coords", default=None, choices=["f32", "f64"], help="coordinates_eltype (default: eltype)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the other precision set-ups")
    ap.add_argument("--shuffle", action="store_true", help="random particle order in the ODE vectors")
    ap.add_argument("--sort", action="store_true", help="tpb_sort_system (SortingCallback) before timing; with --shuffle")
    ap.add_argument("--no-slip", action="store_true", help="no-slip wall (wall viscosity = fluid viscosity) on a "
                    "synthetic velocity field instead of the free-slip headline workload")
    ap.add_argument("--evolve", type=int, default=0, help="time steps to run before measuring (evolved state)")
    ap.add_argument("--per-gpu", type=float, default=12.5e6, help="N > 1: fluid particles per GPU of the "
                    "weak-scaling headline (BASELINE config 4: 12.5 M per GPU = 100 M on 8)")
    ap.add_argument("--summation", action="store_true", help="SummationDensity instead of ContinuityDensity (with --quick)")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (tuning runs)")
    ap.add_argument("--e2e-only", action="store_true", help="host-pointer (e2e) timing only (tuning runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from trixiparticles.jl_b200 import slabs_bench
        slabs_bench.run(args)
        return
    run_single(args)


if __name__ == "__main__":
    main()
