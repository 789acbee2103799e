#!/usr/bin/env python
"""bench.py -- grid-point-steps/s of the spectral RK4 step (FP64) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 16384] [--impl reference]

A "step" is one pass of the hot path `dt = cfl(); advance_rk4()` (forced, hyperviscous;
SURVEY.md 8d config 4 at N=1) over the whole grid. One JSON line is printed by rank 0.

  value     device-resident throughput: K steps of ns_spectral_run_steps in NS_SYNC_LAZY
            mode, state already in HBM, timed with CUDA events on the solver's own stream
            (max over ranks); the per-step forcing box still comes from the host stream.
  e2e       the same metric through the reference-facing C entry points one step per call
            (ns_spectral_run_steps(data, 1, forcing)): every step copies the host-generated
            forcing host->device and reads the step's result scalars (dt, t, step count)
            device->host, wall clock around the calls.
  e2e_eager (extra) the strict drop-in loop `data->dt = ns_spectral_compute_cfl_timestep();
            ns_spectral_advance_rk4()` under NS_SYNC_EAGER: every array the reference
            touches crosses PCIe both ways each step.
  roofline  dominant kernel against the measured HBM peak in MEASURED_PEAKS.json:
            algorithmic bytes per launch (16A per stage, A = 16*nkx*ny bytes, split
            over the kernels of this decomposition) / mean launch time from CUDA events on the
            launching stream (nsb_profile_begin/_end), plus the whole-step figure at 512 B per
            grid-point-step.
  --gpus N  (torchrun, one rank per GPU) the SAME grid slab-decomposed over N GPUs (strong
            scaling): kx-column blocks in spectral space, y-row blocks in physical space, the
            FFT transposes as NVLink peer stores fused into the kernels; ranks exchange CUDA IPC
            handles over the gloo group once, the data path uses no collective library.
  cpu_baseline  oracle/_ref (the unmodified reference sources + shim FFT; FFTW is not
            installable offline) timed on this box's host cores on a bounded sample.

`--impl reference` times that CPU implementation instead (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "grid_point_steps_per_sec"
UNIT = "grid-point-steps/s"
HV = dict(coefficient=5.0, order=4, cutoff_ratio=0.3)
FORCING = dict(kmin=0.5, kmax=4.0, energy_target=1.25, tau=0.4, seed=42)
RE = 1.0e4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synthetic_spectrum_block(kk=64, seed=7, k0=8.0, max_omega=5.0):
    """Random-phase turbulence field, a(k) = k^2 exp(-(k/k0)^2), restricted to
    |kx|,|ky| <= kk where a(k) is non-negligible, Hermitian on kx = 0, scaled to
    max|omega| = max_omega. The block is resolution independent."""
    rng = np.random.default_rng(seed)
    i = np.arange(kk + 1)[None, :]
    b = np.arange(-kk, kk + 1)[:, None]
    k = np.sqrt(i * i + b * b)
    blk = k * k * np.exp(-(k / k0) ** 2) * np.exp(2j * np.pi * rng.random((2 * kk + 1, kk + 1)))
    blk[kk, 0] = 0.0
    col = blk[:, 0].copy()
    blk[:, 0] = 0.5 * (col + np.conj(col[::-1]))
    n = 512
    full = np.zeros((n, n // 2 + 1), dtype=np.complex128)
    full[np.arange(-kk, kk + 1) % n, : kk + 1] = blk
    peak = np.abs(np.fft.irfft2(full, s=(n, n)) * n * n).max()
    return blk * (max_omega / peak), kk


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, flag in zip(names, r[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), \
        int(os.environ.get("WORLD_SIZE", "1"))


def init_dist(world):
    if world <= 1:
        return None
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="gloo")
    return dist


def all_max(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ----------------------------------------------------------------------------- CPU arm

def cpu_reference_run(n, steps, warmup):
    """oracle/_ref (kind 'reference') or the C port, all host threads, forced workload."""
    import orc
    kind = "ref" if orc.have("ref") else "port"
    if not orc.have(kind):
        orc.build_oracles()
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    o = orc.Oracle(kind, n, n, Re=RE)
    o.set_hyperviscosity(1, HV["coefficient"], HV["order"], HV["cutoff_ratio"])
    blk, kk = synthetic_spectrum_block()
    w = o.arr("omega_hat")
    w[...] = 0
    w[np.arange(-kk, kk + 1) % n, : kk + 1] = blk
    o.compute_derivatives(o.sp)
    o.init_forcing("white", kmin=FORCING["kmin"], kmax=FORCING["kmax"], energy_target=FORCING["energy_target"],
                   tau=FORCING["tau"], seed=FORCING["seed"])
    for _ in range(warmup):
        o.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step()
    dt = time.perf_counter() - t0
    o.close()
    label = ("reference solver (unmodified /root/reference sources) + shim FFT (FFTW unavailable offline)"
             if kind == "ref" else "C port of the reference + own FFT")
    return {"value": n * n * steps / dt, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port",
            "sample": f"{n}x{n} forced turbulence, {steps} steps after {warmup} warm-up; {label}",
            "seconds": dt, "ms_per_step": 1e3 * dt / steps}


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    n = args.cpu_n
    steps = max(1, min(args.steps, args.cpu_steps))
    warm = max(0, min(args.warmup, 1))
    r = cpu_reference_run(n, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"forced 2-D turbulence {args.n}x{args.n} FP64 (config 4); timed on a bounded "
                                   f"{n}x{n} sample of it", "Re": RE, "hyperviscosity": HV, "forcing": "white"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

def run_gpu_arm(args):
    import __graft_entry__ as ge
    import hostlib
    import orc
    rank, local_rank, world = dist_env()
    dist = init_dist(world)
    nsb = ge.load_package()
    nsblib = nsb.load_nsb()                      # raises if missing: no fallback
    host = hostlib.load(hostlib.PRODUCT)
    if nsblib.nsb_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device visible and there is no CPU fallback")
    n = args.n
    device = local_rank % max(1, nsblib.nsb_device_count())
    os.environ["NS_SPECTRAL_DEVICE"] = str(device)

    s = hostlib.Solver(host, n, n, Re=RE)
    host.ns_spectral_backend_context.restype = C.c_void_p
    host.ns_spectral_backend_context.argtypes = [C.c_void_p]
    ctx = C.c_void_p(host.ns_spectral_backend_context(s.sp))
    if world > 1:
        # slab decomposition: partition, then swap the CUDA IPC handles of T, T' and the scalar block
        import torch
        nsblib.nsb_mgpu_setup.argtypes = [C.c_void_p, C.c_int, C.c_int]
        nsblib.nsb_mgpu_export.argtypes = [C.c_void_p, C.c_char_p]
        nsblib.nsb_mgpu_import.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        if nsblib.nsb_mgpu_setup(ctx, rank, world) != 0:
            raise SystemExit("bench.py: " + nsblib.nsb_last_error().decode())
        buf = C.create_string_buffer(192)
        if nsblib.nsb_mgpu_export(ctx, buf) != 0:
            raise SystemExit("bench.py: " + nsblib.nsb_last_error().decode())
        mine = torch.tensor(list(buf.raw), dtype=torch.uint8)
        gathered = [torch.zeros(192, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, mine)
        for peer in range(world):
            if peer != rank and nsblib.nsb_mgpu_import(ctx, peer, bytes(gathered[peer].tolist())) != 0:
                raise SystemExit("bench.py: " + nsblib.nsb_last_error().decode())
        dist.barrier()
    blk, kk = synthetic_spectrum_block()
    w = s.arr("omega_hat")
    w[np.arange(-kk, kk + 1) % n, : kk + 1] = blk
    h = orc.Hyperviscosity()
    host.ns_spectral_hyperviscosity_default(C.byref(h))
    h.enabled, h.coefficient, h.order, h.cutoff_ratio = 1, HV["coefficient"], HV["order"], HV["cutoff_ratio"]
    host.ns_spectral_hyperviscosity_set(s.sp, C.byref(h))
    cfg = orc.ForcingConfig()
    host.ns_forcing_config_default(C.byref(cfg))
    cfg.kmin, cfg.kmax, cfg.energy_target = FORCING["kmin"], FORCING["kmax"], FORCING["energy_target"]
    cfg.correlation_time, cfg.seed, cfg.type = FORCING["tau"], FORCING["seed"], 0  # every rank replays the same stream
    fs = orc.ForcingState()
    assert host.ns_forcing_state_init(C.byref(fs), C.byref(cfg), s.s.k2, s.s.nkx, s.s.nky) == 0
    ib, jb = C.c_int(), C.c_int()
    host.ns_forcing_box_extent(C.byref(fs), C.byref(ib), C.byref(jb))
    force_bytes = (2 * jb.value + 1) * (ib.value + 1) * 16

    host.ns_spectral_set_sync_policy(s.sp, 1)          # LAZY: omega_hat now lives in HBM
    for _ in range(args.warmup):
        assert host.ns_spectral_run_steps(s.sp, 1, C.byref(fs)) == 0

    # ---- device-resident timed region: exactly K steps
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    if dist is not None:
        dist.barrier()
    nsblib.nsb_sync(ctx)
    l0 = nsblib.nsb_launch_count(ctx)
    nsblib.nsb_timer_start(ctx)
    assert host.ns_spectral_run_steps(s.sp, args.steps, C.byref(fs)) == 0
    ms = C.c_float()
    nsblib.nsb_timer_stop(ctx, C.byref(ms))
    nsblib.nsb_sync(ctx)
    if dist is not None:
        dist.barrier()
    launches = nsblib.nsb_launch_count(ctx) - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = all_max(dist, ms.value / args.steps)
    total_points = float(n) * n                       # ONE grid, slab-decomposed over the ranks
    value = total_points / (ms_step * 1e-3)

    # ---- e2e: one step per call through the C entry points, forcing H2D + scalars D2H each step
    e2e_steps = max(1, min(args.steps, 10))
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert host.ns_spectral_run_steps(s.sp, 1, C.byref(fs)) == 0
    nsblib.nsb_sync(ctx)
    e2e_s = all_max(dist, (time.perf_counter() - t0) / e2e_steps)
    e2e = {"value": total_points / e2e_s, "unit": UNIT, "h2d_bytes_per_step": force_bytes,
           "d2h_bytes_per_step": 2112, "steps": e2e_steps,
           "api": "ns_spectral_run_steps(data, 1, forcing) per step, NS_SYNC_LAZY"}

    # ---- per-kernel timing for the roofline (separate short pass, CUDA events per launch)
    nsblib.nsb_profile_begin(ctx)
    prof_steps = 2
    assert host.ns_spectral_run_steps(s.sp, prof_steps, C.byref(fs)) == 0
    kms = (C.c_double * 6)()
    kcnt = (C.c_int * 6)()
    nsblib.nsb_profile_end(ctx, kms, kcnt)
    names = ["k_alpha1024<inv> (stage-1 spectral ops + inverse y FFT)", "k_chain_B_tma", "k_chain_G",
             "k_chain_F", "k_alpha1024<fwd,inv> (forward y FFT + RK update + next inverse y FFT)", "other"]
    if n <= 1024:
        names = ["k_alpha1024<inv>", "k_beta_inv (unused)", "k_gamma1024", "k_beta_fwd (unused)",
                 "k_alpha1024<fwd,inv>", "other"]
    per_kernel = {nm: {"launches_per_step": int(kcnt[k]) // prof_steps,
                       "ms_per_launch": (kms[k] / kcnt[k]) if kcnt[k] else None,
                       "ms_per_step": kms[k] / prof_steps} for k, nm in enumerate(names)}
    A = 16.0 * (n // 2 + 1) * n
    peak, peak_src = measured_peaks()
    top_i = max(range(5), key=lambda k: kms[k])
    top = names[top_i]
    # SURVEY 8d contract: 16A per stage = 64A per step, A = 16*nkx*ny bytes (full, unpruned arrays):
    # (i) spectral ops + inverse y pass 1A in / 4A out; (ii) inverse x + Jacobian + forward x 4A in /
    # 1A out; (iii) forward y pass + RK 4A in / 2A out. Per step and kernel kind of this decomposition:
    # stage-1 alpha_inv = (i) = 5A; fused alpha = 3 x [(iii)+(i)] + 1 x (iii) = 39A; G = 4 x (ii) = 20A.
    # B and F are passes the model does not grant; they are charged at their own in+out arrays
    # (4 x 8A, 4 x 2A) for reference only. Each rank moves 1/world of these bytes.
    alg_step = [5 * A, 32 * A, 20 * A, 8 * A, 39 * A]
    top_launches = max(1, int(kcnt[top_i]) // prof_steps)
    alg_launch = alg_step[top_i] / top_launches / world
    tl = per_kernel[top]["ms_per_launch"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = json.load(open(tpath)).get(str(n), {}).get(str(top_i))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": top, "achieved": alg_launch / (tl * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": alg_launch / (tl * 1e-3) / 1e9 / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_launch,
                "ms_per_launch": tl,
                "step": {"algorithmic_bytes_per_grid_point_step": 512,
                         "achieved": 512.0 * (float(n) * n / (ms_step * 1e-3)) / 1e9 / world,
                         "frac": 512.0 * (float(n) * n / (ms_step * 1e-3)) / 1e9 / world / peak,
                         "note": "per GPU: 512 B x grid-point-steps/s / n_gpus against one GPU's HBM peak"},
                "per_kernel": per_kernel,
                "per_kernel_note": "CUDA events around every launch need one stream: measured in a separate pass "
                                   "with the chain in ONE slab range; in the timed region B, G, F run as up to 16 "
                                   "slab ranges on two streams (profiles/README.md has the ncu launch list)"}
    if world > 1:
        # NVLink: every stage pushes 4 T fields + 1 T' field of the retained columns; (world-1)/world of
        # a rank's share leaves the GPU. SURVEY 8e contract figure (unpruned) alongside the real bytes.
        ncol = (2 * (n // 2 + 1)) // 3 + 1
        real = 4 * 5 * (16.0 * ncol * n / world) * (world - 1) / world
        contract = 20 * (A / world) * (world - 1) / world
        roofline["nvlink"] = {"bytes_sent_per_gpu_per_step": real, "contract_bytes_per_gpu_per_step": contract,
                              "achieved_GBs": real / (ms_step * 1e-3) / 1e9,
                              "contract_GBs": contract / (ms_step * 1e-3) / 1e9,
                              "peak_GBs": 770.0, "peak_source": "measured peer copy per direction (B200_PROFILING.md)",
                              "frac": contract / (ms_step * 1e-3) / 1e9 / 770.0}

    # fingerprint of the whole run: t and dt after all steps depend on max|u| over the whole field at every
    # step, so runs with the same --warmup/--steps must agree bit for bit across GPU counts
    fingerprint = {"t": float(s.s.t).hex(), "dt": float(s.s.dt).hex(), "t_dec": s.s.t, "dt_dec": s.s.dt,
                   "steps_taken": args.warmup + args.steps + e2e_steps + prof_steps}

    # ---- strict drop-in loop (EAGER): bounded, single GPU only, small step count
    e2e_eager = None
    if world == 1 and args.eager_steps > 0:
        host.ns_spectral_set_sync_policy(s.sp, 0)
        t0 = time.perf_counter()
        for _ in range(args.eager_steps):
            s.s.dt = host.ns_spectral_compute_cfl_timestep(s.sp)
            host.ns_spectral_advance_rk4(s.sp)
        te = (time.perf_counter() - t0) / args.eager_steps
        m_bytes, n_bytes = 16 * (n // 2 + 1) * n, 8 * n * n
        e2e_eager = {"value": float(n) * n / te, "unit": UNIT, "steps": args.eager_steps,
                     "h2d_bytes_per_step": m_bytes + 2 * n_bytes, "d2h_bytes_per_step": 4 * m_bytes + 3 * n_bytes,
                     "api": "ns_spectral_compute_cfl_timestep + ns_spectral_advance_rk4, NS_SYNC_EAGER (pageable host arrays)"}

    host.ns_forcing_state_free(C.byref(fs))
    plan = (C.c_int * 8)()
    nsblib.nsb_get_plan(ctx, plan)
    s.close()

    if rank != 0:
        return
    cpu = None
    if args.cpu_steps > 0:
        r = cpu_reference_run(args.cpu_n, args.cpu_steps, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"forced 2-D turbulence {n}x{n} FP64, hyperviscosity + white forcing, adaptive CFL "
                                   f"dt (BASELINE config {'4' if n == 16384 else '3' if n == 4096 else '5 grid, forced' if n == 32768 else 'n=' + str(n)})",
                       "Re": RE, "hyperviscosity": HV, "forcing": {"type": "white", **FORCING},
                       "init": "random-phase spectrum k^2 exp(-(k/8)^2), max|omega| = 5",
                       "parallelism": "1 GPU" if world == 1 else
                       f"slab decomposition over {world} GPUs: kx-column blocks / y-row blocks, FFT transposes as "
                       "NVLink peer stores fused into the kernels (one process per GPU, CUDA IPC)",
                       "plan": {"Y1": plan[0], "Y2": plan[1], "X1": plan[2], "X2": plan[3]},
                       "l2": "inputs larger than L2: working set >> 126 MB at this size, no flush needed"
                             if n >= 2048 else "working set fits L2 at this size"},
            "clocks": clocks, "e2e": e2e, "e2e_eager": e2e_eager, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "fingerprint": fingerprint}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--grid", dest="n", type=int, default=16384,
                    help="grid size (nx = ny); under torchrun use --grid (its parser takes --n for its own options)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-n", type=int, default=2048, help="grid of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed CPU steps (0 = skip cpu_baseline)")
    ap.add_argument("--eager-steps", type=int, default=1, help="strict EAGER drop-in steps (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
