#!/usr/bin/env python
"""bench.py — surface-flux throughput of the coflux CUDA path (and of the CPU oracle beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl coflux|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json config 4 / north_star target): one `update_state!` flux solve (atmosphere →
ocean-grid interpolation + Monin–Obukhov similarity solve + net ocean flux assembly) per step on the
synthetic 1/12° global lat-lon grid 4320×1800×75, Float64, `:default` flux configuration, ocean-only
(prescribed atmosphere).  N GPUs split the grid into longitude slabs (strong scaling, zero-message
halo-ring mode).  One JSON line is printed by rank 0.  metric = Mcells/s = Nx·Ny / t_step / 1e6.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NX, NY, NZ = 4320, 1800, 75                  # 1/12° (config 4)
QNX, QNY, QNZ = 1440, 600, 10                # 1/4°  (config 2), reported as an extra
QUERY_TIME = 1.37 * 3 * 3600.0
WORDS_STEP = 29                              # SURVEY §8d fused-minimum words per cell for the flux solve
WORDS_FLUX_KERNEL = 26                       # share of the flux kernel (no land mask in the bench): fi fj + ocean u v T S
                                             #   + 8 exchange + 6 interface + 6 tracer/radiative outputs
WORDS_STRESS_KERNEL = 2                      # τx τy (its reads of ρτx ρτy are re-reads of intermediates)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


TRAFFIC_FILE = "profiles/r02_traffic.json"


def load_traffic(cells):
    """dram__bytes_read.sum + dram__bytes_write.sum of the flux kernel.  NOT measured in this run (a bench number is never
    taken under a profiler): it comes from the committed `ncu --set full` capture of the same kernel on the same 1/12° grid
    (bytes per cell in profiles/r02_traffic.json) scaled to this launch; `traffic_source` in the JSON line says so."""
    p = os.path.join(ROOT, TRAFFIC_FILE)
    try:
        return float(json.load(open(p))["flux_tile_kernel_f64_dram_bytes_per_cell"]) * cells
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_host_case(Nx_global, Ny, bits, rank, world):
    import climaocean.jl_b200 as cj
    dtype = np.float64 if bits == 64 else np.float32
    full = cj.LatitudeLongitudeGrid((Nx_global, Ny, 1), latitude=(-75.0, 75.0), halo=(7, 7, 0), dtype=dtype)
    grid = full.slab(rank, world) if world > 1 else full
    host = cj.SurfaceFluxData.synthetic(grid, ring=1)
    return grid, host


def make_cfg(grid, Nz, bits, device_index, flux_configuration="default"):
    import climaocean.jl_b200 as cj
    cfg = cj.default_config(grid.Nx, grid.Ny, Nz, bits, flux_configuration)
    cfg.device = device_index
    cfg.grid.ring = 1            # zero-message mode: fluxes are computed into one halo ring
    return cfg


def barrier(dist):
    if dist is not None:
        dist.barrier()


def gather_over_ranks(dist, x, device, world):
    """x of every rank, on every rank (list of floats)."""
    if dist is None:
        return [float(x)]
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def bit_checksum(field, x0=None, x1=None):
    """Order-independent checksum of the interior of a 2-D output Field: the sum of the raw bit patterns (mod 2⁶⁴) of the
    columns [x0, x1) — equal checksums ⇔ (to all intents) bit-identical data, and slab checksums add up to the global one."""
    import torch
    Hx, Hy, _ = field.halo
    a = field.data[0, Hy:field.data.shape[1] - Hy, Hx:field.data.shape[2] - Hx]
    if x0 is not None:
        a = a[:, x0:x1]
    bits = a.contiguous().view(torch.int64 if a.dtype == torch.float64 else torch.int32).to(torch.int64)
    return int(bits.sum().item())


CHECK_FIELDS = (("net", "u"), ("net", "v"), ("net", "T"), ("net", "S"), ("ao", "latent_heat"), ("ao", "sensible_heat"))


def max_over_ranks(dist, x, device):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_device_steps(eng, dev, steps, warmup, dist, device, sampler=None):
    """K update_state! calls on device-resident inputs, CUDA events on the launching stream."""
    import torch
    inp, out = dev.update_bundles()
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        eng.update_state(inp, out, QUERY_TIME, stream)
    torch.cuda.synchronize()
    eng.profile(True)
    eng.profile_read()
    l0 = eng.launches
    barrier(dist)
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        eng.update_state(inp, out, QUERY_TIME, stream)
    e1.record(stream)
    torch.cuda.synchronize()
    barrier(dist)
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1) / steps
    flux_ms, stress_ms, calls = eng.profile_read()
    eng.profile(False)
    launches = eng.launches - l0
    time_device_steps.local_ms = ms                       # this rank's own time (the return value is the max over ranks)
    ms = max_over_ranks(dist, ms, device)
    return ms, launches, flux_ms / max(calls, 1), stress_ms / max(calls, 1), clocks


def time_e2e_steps(eng, dev, host, grid, bits, steps, warmup, dist, device):
    """K coflux_update_state_host calls: per step H2D of the ocean surface planes from pinned memory, the
    two kernels, D2H of the four net fluxes into pinned memory; wall clock around the blocking calls."""
    import torch
    from climaocean.jl_b200 import _abi
    H = grid.halo[0]
    planes = {n: torch.from_numpy(np.ascontiguousarray(host.ocean[n].data[0])).pin_memory() for n in ("u", "v", "T", "S")}
    outs = {n: torch.empty_like(planes["u"]).pin_memory() for n in ("u", "v", "T", "S")}
    step = _abi.HostStep(planes["u"].data_ptr(), planes["v"].data_ptr(), planes["T"].data_ptr(), planes["S"].data_ptr(),
                         outs["u"].data_ptr(), outs["v"].data_ptr(), outs["T"].data_ptr(), outs["S"].data_ptr(), None, None, H, 0)
    series = dev.atmos_series()
    h2d = d2h = 0
    for _ in range(max(warmup, 1)):
        h2d, d2h = eng.update_state_host(series, step, QUERY_TIME)
    barrier(dist)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.update_state_host(series, step, QUERY_TIME)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3 / steps
    barrier(dist)
    dt = max_over_ranks(dist, dt, device)
    checksum = float(outs["T"].double().abs().sum())
    return dt, h2d, d2h, checksum


def oracle_rate(host, cfg_full, rows, threads):
    """Time the CPU oracle's update_state on the first `rows` latitude rows of the workload."""
    from oracle import pyoracle
    from climaocean.jl_b200 import _abi
    cfg = _abi.Config.from_buffer_copy(cfg_full)
    cfg.grid.Ny = rows
    cfg.grid.Nz = 1
    pyoracle.set_threads(threads)
    inp, out = host.update_bundles()
    t0 = time.perf_counter()
    pyoracle.update_state(cfg, inp, out, QUERY_TIME)
    dt = time.perf_counter() - t0
    return dt, cfg.grid.Nx * rows


def cpu_baseline(host, cfg, target_seconds=12.0):
    """Oracle port on all host cores, on a bounded sample of the same workload (≈ target_seconds of CPU wall
    time): as many latitude rows as fit, the whole grid repeated when one pass is shorter than the target.
    Uses the CPU-baseline build of the oracle (-O3 -march=native, compiled here for this host: oracle/Makefile `fast`)."""
    from oracle import pyoracle
    pyoracle.use_fast_build()
    cores = os.cpu_count() or 1
    pyoracle.set_threads(cores)
    oracle_rate(host, cfg, 4, cores)                                   # spin up the thread team
    Ny, Nx = host.grid.Ny, host.grid.Nx
    rows = 16
    dt, cells = oracle_rate(host, cfg, rows, cores)
    while dt < 1.0 and rows < Ny:                                      # grow the probe until it is long enough to trust
        rows = min(Ny, rows * 4)
        dt, cells = oracle_rate(host, cfg, rows, cores)
    rate = cells / dt
    rows = int(min(Ny, max(rows, rate * target_seconds / Nx)))
    total_t, total_c, reps = 0.0, 0, 0
    while total_t < target_seconds * 0.8 and reps < 16:
        dt, cells = oracle_rate(host, cfg, rows, cores)
        total_t += dt; total_c += cells; reps += 1
    # one core beside it: ≈ 2 s of the same rows
    rows1 = int(max(4, min(Ny, rate / max(cores, 1) * 2.0 / Nx)))
    dt1, cells1 = oracle_rate(host, cfg, rows1, 1)
    pyoracle.set_threads(cores)
    return {"value": total_c / total_t / 1e6, "unit": "Mcells/s", "cores": cores, "kind": "port",
            "single_core_value": cells1 / dt1 / 1e6,
            "build": "gcc -O3 -march=native -fopenmp (oracle/Makefile: fast), built on this host",
            "sample": f"CPU oracle (C restatement, OpenMP) update_state on the first {rows} of {Ny} latitude rows "
                      f"({Nx * rows} cells) of the same workload, {reps} pass(es), {total_t:.1f} s",
            "note": "the Julia reference cannot run here (no Julia, dependency un-vendored); this is the oracle port"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["COFLUX_NO_LIBRARY"] = "1"       # this arm must not load the product: grid / synthetic-input helpers only
    from oracle import pyoracle
    from climaocean.jl_b200 import _abi
    pyoracle.use_fast_build()                    # -O3 -march=native build of the oracle, compiled here for this host
    grid, host = make_host_case(NX, NY, 64, 0, 1)
    cfg = pyoracle.default_config(_abi.Config(), grid.Nx, grid.Ny, 1, 64, "default")     # defaults stated on the oracle side
    cfg.grid.ring = 1
    cores = os.cpu_count() or 1
    pyoracle.set_threads(cores)
    oracle_rate(host, cfg, 4, cores)
    rows = 16
    dt, cells = oracle_rate(host, cfg, rows, cores)
    while dt < 0.5 and rows < NY:
        rows = min(NY, rows * 4)
        dt, cells = oracle_rate(host, cfg, rows, cores)
    rows = int(min(NY, max(16, (cells / dt) * 2.0 / NX)))              # ≈ 2 s per step
    for _ in range(args.warmup):
        oracle_rate(host, cfg, rows, cores)
    t0 = time.perf_counter()
    cells = 0
    for _ in range(args.steps):
        _, c = oracle_rate(host, cfg, rows, cores)
        cells += c
    dt = time.perf_counter() - t0
    v = cells / dt / 1e6
    sample = f"{rows} of {NY} latitude rows ({NX * rows} cells) of the 1/12° workload per step"
    print(json.dumps({
        "impl": "reference", "metric": "surface_flux_mcells_per_s", "value": v, "unit": "Mcells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": v, "unit": "Mcells/s", "cores": cores, "kind": "port", "sample": sample,
                         "build": "gcc -O3 -march=native -fopenmp (oracle/Makefile: fast), built on this host"},
        "e2e": {"value": v, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port on host cores; the Julia reference is not runnable in this environment (DESIGN.md §3)"}))


def workload_config(world):
    return {"workload": f"1/12deg global LatitudeLongitudeGrid {NX}x{NY}x{NZ}, synthetic PrescribedAtmosphere 640x320x8, "
                        "ocean-only update_state! flux solve (interpolate + similarity solve + net ocean flux assembly)",
            "flux_configuration": "default (Edson psi, constant Charnock 0.02, convergence 1e-8 / maxiter 100)",
            "decomposition": f"{world} longitude slab(s), zero-message halo-ring mode",
            "l2": "inputs+outputs per step (1.8 GB) exceed the 126 MB L2; no flush needed",
            "ocean_parents": "full 3-D (Nz+14 levels) device arrays; the path reads the k=Nz-1 plane"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="coflux", choices=["coflux", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the F32 / quarter-degree / ice-ocean extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import climaocean.jl_b200 as cj
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: coflux has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=torch.device(device))
        dist = dist_mod
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    peak, peak_src = load_peaks()

    # ---- headline: 1/12°, Float64, :default ----
    grid, host = make_host_case(NX, NY, 64, rank, world)
    dev = host.to_device_columns(device, NZ)
    cfg = make_cfg(dev.grid, NZ, 64, local)
    eng = cj.Engine(cfg)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, flux_ms, stress_ms, clocks = time_device_steps(eng, dev, args.steps, args.warmup, dist, device, sampler)
    cells_global = NX * NY
    cells_local = grid.Nx * grid.Ny
    value = cells_global / (ms * 1e-3) / 1e6
    rank_kernel_ms = gather_over_ranks(dist, flux_ms, device, world)
    rank_step_ms = gather_over_ranks(dist, time_device_steps.local_ms, device, world)
    its = dev.iterations.numpy()[0, 7:-7, 7:-7]
    roof = {"bound": "hbm", "kernel": "flux_tile_kernel<double,1,1,1280,1>, 320 threads x 2 CTAs/SM (fused interpolate + similarity solve + tracer/radiative assembly)",
            "achieved": cells_local * WORDS_FLUX_KERNEL * 8 / (flux_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "peak_source": peak_src, "traffic": load_traffic(cells_local),
            "algorithmic_bytes_per_cell": WORDS_FLUX_KERNEL * 8, "kernel_ms": flux_ms, "stress_kernel_ms": stress_ms,
            "step_achieved_29_words": cells_local * WORDS_STEP * 8 / (ms * 1e-3) / 1e9,
            "note": "the converged Float64 solve is bound by dependent FP64 latency / issue, not by HBM (DESIGN.md §4, profiles/README.md)",
            "iterations_mean": float(its.mean()), "iterations_max": int(its.max())}
    roof["frac"] = roof["achieved"] / peak
    roof["step_frac_29_words"] = roof["step_achieved_29_words"] / peak

    # ---- end to end through the HOST-buffer C-ABI entry ----
    cfg_h = make_cfg(grid, 1, 64, local)
    eng_h = cj.Engine(cfg_h)
    e2e_ms, h2d, d2h, _ = time_e2e_steps(eng_h, dev, host, grid, 64, max(3, args.steps // 4), 2, dist, device)
    e2e = {"value": cells_global / (e2e_ms * 1e-3) / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms, "api": "coflux_update_state_host (pinned host planes in, pinned host net fluxes out)"}
    eng_h.close()

    extras = {}
    if not args.no_extras and world == 1:
        # Float32 on the same grid
        g32, h32 = make_host_case(NX, NY, 32, 0, 1)
        d32 = h32.to_device_columns(device, NZ)
        e32 = cj.Engine(make_cfg(d32.grid, NZ, 32, local))
        m32, _, f32ms, s32ms, _ = time_device_steps(e32, d32, max(5, args.steps // 2), 3, None, device)
        it32 = d32.iterations.numpy()[0, 7:-7, 7:-7]
        extras["f32"] = {"value": cells_global / (m32 * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": m32, "flux_kernel_ms": f32ms,
                         "roofline_frac": cells_global * WORDS_FLUX_KERNEL * 4 / (f32ms * 1e-3) / 1e9 / peak,
                         "iterations_mean": float(it32.mean()), "iterations_max": int(it32.max())}
        e32.close(); del d32, h32
        # the other OMIP flux configurations, Float64
        for name in ("corrected", "ncar"):
            ec = cj.Engine(make_cfg(dev.grid, NZ, 64, local, name))
            mc, _, fc, sc, _ = time_device_steps(ec, dev, max(5, args.steps // 2), 3, None, device)
            extras[name] = {"value": cells_global / (mc * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": mc, "flux_kernel_ms": fc,
                            "roofline_frac": cells_global * WORDS_FLUX_KERNEL * 8 / (fc * 1e-3) / 1e9 / peak}
            ec.close()
        # 1/4° (config 2): smaller than the 1/12° grid; flush L2 between steps is unnecessary (200 MB/step > 126 MB L2)
        gq, hq = make_host_case(QNX, QNY, 64, 0, 1)
        dq = hq.to_device_columns(device, QNZ)
        eq = cj.Engine(make_cfg(dq.grid, QNZ, 64, local))
        mq, _, fq, sq, _ = time_device_steps(eq, dq, max(5, args.steps), 3, None, device)
        extras["quarter_degree_1440x600x10_f64"] = {"value": QNX * QNY / (mq * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": mq,
                                                    "flux_kernel_ms": fq, "stress_kernel_ms": sq}
        eq.close(); del dq, hq
        # NormalizeSalinity (omip_simulation.jl:187-220) on the net salinity flux of the headline run: HBM bound, the
        # flux plane is read twice (sums, subtraction) and written once = 3 words per cell
        st0 = torch.cuda.current_stream()
        norm = dev.salinity_normalization()
        tn = []
        for k in range(8):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(st0); eng.normalize_salinity_flux(norm, st0); b_.record(st0)
            torch.cuda.synchronize()
            if k >= 3:
                tn.append(a_.elapsed_time(b_))
        tnm = float(np.mean(tn))
        n0 = eng.launches
        eng.normalize_salinity_flux(norm, st0)
        extras["normalize_salinity_f64"] = {"ms": tnm, "launches": eng.launches - n0, "algorithmic_bytes_per_cell": 24,
                                            "achieved_GBs": cells_global * 24 / (tnm * 1e-3) / 1e9,
                                            "roofline_frac": cells_global * 24 / (tnm * 1e-3) / 1e9 / peak,
                                            "note": "one cooperative launch: sums, grid barrier, subtraction (the second pass reads the flux plane from L2)"}
        # running time averages (omip_diagnostics.jl:125-158) attached to the step: epilogues of the flux and stress kernels,
        # 6 ocean-only averages read + written per cell = 12 more words; no extra launch
        dev.allocate_averages()
        eng.attach_flux_averages(dev.flux_averages(0.0, 600.0))
        mavg, lavg, favg, savg, _ = time_device_steps(eng, dev, max(5, args.steps // 2), 3, None, device)
        eng.attach_flux_averages(None)
        extras["step_with_time_averages_f64"] = {"ms_per_step": mavg, "flux_kernel_ms": favg, "stress_kernel_ms": savg,
                                                 "extra_ms_vs_plain_step": mavg - ms, "launches_per_step": lavg / max(5, args.steps // 2)}
        del dev.averages
        # closure surface-forcing front ends (KPP u★, Bo; NEMO-TKE u★², e_surf): stand-alone kernel, 6 reads + 4 writes per cell
        cf = dev.closure_forcing()
        netb = dev.net_ocean_fluxes()
        tc = []
        for k in range(8):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(st0); eng.closure_surface_forcing(netb, cf, st0); b_.record(st0)
            torch.cuda.synchronize()
            if k >= 3:
                tc.append(a_.elapsed_time(b_))
        tcm = float(np.mean(tc))
        extras["closure_surface_forcing_f64"] = {"ms": tcm, "algorithmic_bytes_per_cell": 80,
                                                 "achieved_GBs": cells_global * 80 / (tcm * 1e-3) / 1e9,
                                                 "roofline_frac": cells_global * 80 / (tcm * 1e-3) / 1e9 / peak}
        # sea-ice–ocean kernel (HBM bound: 2·Nz + 13 words per column), 1/12°, Nz = 75
        gi = cj.LatitudeLongitudeGrid((NX, NY, 1), latitude=(-75.0, 75.0), halo=(7, 7, 0))
        hi = cj.SurfaceFluxData.synthetic(gi, with_ice=True)
        di = hi.to_device_columns(device, NZ, fill_columns=True)
        ei = cj.Engine(make_cfg(di.grid, NZ, 64, local))
        cols, ice, io = di.ocean_columns(), di.sea_ice_state(), di.ice_ocean_fluxes()
        T0 = di.ocean["T"].data.clone()
        st = torch.cuda.current_stream()
        tms = []
        for k in range(6):
            di.ocean["T"].data.copy_(T0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); ei.compute_sea_ice_ocean_fluxes(cols, ice, 600.0, io, st); b.record(st)
            torch.cuda.synchronize()
            if k >= 2:
                tms.append(a.elapsed_time(b))
        tio = float(np.mean(tms))
        words = 2 * NZ + 13
        extras["sea_ice_ocean_kernel_f64"] = {"ms": tio, "Mcells/s": cells_global / (tio * 1e-3) / 1e6,
                                              "algorithmic_bytes_per_cell": words * 8,
                                              "achieved_GBs": cells_global * words * 8 / (tio * 1e-3) / 1e9,
                                              "roofline_frac": cells_global * words * 8 / (tio * 1e-3) / 1e9 / peak}
        # atmosphere–sea-ice solve with skin temperature (row a7; one cell per thread, SHEBA/Paulson ψ, fixed roughness)
        xch_i, oc_i, ai_i = di.exchange_state(), di.ocean_surface(), di.interface_fluxes("ai")
        ei.interpolate_atmosphere_state(di.atmos_series(), QUERY_TIME, xch_i, st)
        Ttop0 = di.ice["top_temperature"].data.clone()
        tai = []
        for k in range(6):
            di.ice["top_temperature"].data.copy_(Ttop0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); ei.compute_atmosphere_sea_ice_fluxes(xch_i, oc_i, ice, ai_i, st); b.record(st)
            torch.cuda.synchronize()
            if k >= 2:
                tai.append(a.elapsed_time(b))
        tam = float(np.mean(tai))
        extras["atmosphere_sea_ice_kernel_f64"] = {"ms": tam, "Mcells/s": cells_global / (tam * 1e-3) / 1e6,
                                                   "ice_covered_fraction": float((di.ice["concentration"].data > 0).double().mean())}
        # compute_net_sea_ice_fluxes! (top / bottom heat fluxes + face stresses over ice; 12 reads + 4 writes per cell)
        net_i, io_i = di.net_sea_ice_fluxes(), di.ice_ocean_fluxes()
        tni = []
        for k in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); ei.compute_net_sea_ice_fluxes(xch_i, oc_i, ice, ai_i, io_i, net_i, st); b.record(st)
            torch.cuda.synchronize()
            if k >= 3:
                tni.append(a.elapsed_time(b))
        tnim = float(np.mean(tni))
        extras["net_sea_ice_fluxes_kernel_f64"] = {"ms": tnim, "algorithmic_bytes_per_cell": 16 * 8 + 1,
                                                   "achieved_GBs": cells_global * 129 / (tnim * 1e-3) / 1e9,
                                                   "roofline_frac": cells_global * 129 / (tnim * 1e-3) / 1e9 / peak}
        ei.close(); del di, hi, T0

    multi_gpu_check = None
    if world > 1:
        # correctness of the slab decomposition, carried by the scaling run itself: every rank checksums the bit patterns of
        # its slab's outputs; rank 0 solves the WHOLE grid once on its own GPU and checksums the same column ranges.
        import torch
        mine = torch.tensor([bit_checksum(getattr(dev, g)[n]) for g, n in CHECK_FIELDS], dtype=torch.int64, device=device)
        allsums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
        if rank == 0:
            gf, hf = make_host_case(NX, NY, 64, 0, 1)
            df = hf.to_device_columns(device, NZ)
            ef = cj.Engine(make_cfg(df.grid, NZ, 64, local))
            inp_f, out_f = df.update_bundles()
            ef.update_state(inp_f, out_f, QUERY_TIME, torch.cuda.current_stream())
            torch.cuda.synchronize()
            nx = NX // world
            bad, bad_fields = [], {}
            for r in range(world):
                want = [bit_checksum(getattr(df, g)[n], r * nx, (r + 1) * nx) for g, n in CHECK_FIELDS]
                got = [int(v) for v in allsums[r].tolist()]
                if want != got:
                    bad.append(r)
                    bad_fields[str(r)] = [f"{g}.{n}" for (g, n), w_, g_ in zip(CHECK_FIELDS, want, got) if w_ != g_]
            multi_gpu_check = {"bitwise_match_vs_single_gpu": not bad, "mismatching_slabs": bad, "mismatching_fields": bad_fields,
                               "fields": [f"{g}.{n}" for g, n in CHECK_FIELDS],
                               "how": "sum of raw bit patterns (mod 2^64) of every slab's interior vs the same columns of a one-GPU solve of the whole grid"}
            ef.close(); del df, hf
        barrier(dist)
        # mode B: ring = 0, the flux kernel pushes the seam column of ρτx into the east neighbour over NVLink
        from climaocean.jl_b200 import slabs
        gs, hs = make_host_case(NX, NY, 64, rank, world)
        hs0 = cj.SurfaceFluxData.synthetic(gs, ring=0)
        ds = hs0.to_device_columns(device, NZ)
        cfg_s = make_cfg(ds.grid, NZ, 64, local)
        cfg_s.grid.ring = 0
        cfg_s.grid.periodic_x = 0
        es_ = cj.Engine(cfg_s)
        slabs.attach_seam(es_, dist, rank, world)
        ms_s, _, fs_, ss_, _ = time_device_steps(es_, ds, args.steps, args.warmup, dist, device)
        extras["seam_push_mode"] = {"value": cells_global / (ms_s * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms_s,
                                    "flux_kernel_ms": fs_, "stress_kernel_ms_incl_seam_wait": ss_,
                                    "seam_bytes_per_step_per_rank": NY * 8,
                                    "note": "NVLink peer store from inside the flux kernel + stream write/wait-value; no NCCL call per step"}
        barrier(dist)
        es_.seam_detach()
        es_.close()
        del ds, hs0

    cpu = None
    if rank == 0 and world == 1:
        cpu = cpu_baseline(host, make_cfg(grid, 1, 64, 0))

    if rank == 0:
        line = {"metric": "surface_flux_mcells_per_s", "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(world), "roofline": roof, "cpu_baseline": cpu,
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "extras": extras,
                "ranks": {"flux_kernel_ms": {"min": min(rank_kernel_ms), "mean": float(np.mean(rank_kernel_ms)), "max": max(rank_kernel_ms)},
                          "step_ms": {"min": min(rank_step_ms), "mean": float(np.mean(rank_step_ms)), "max": max(rank_step_ms)}},
                "multi_gpu_check": multi_gpu_check,
                "parity_note": "oracle-relative (the Julia reference cannot run here; parity unpinned, DESIGN.md §3)"}
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
