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
  roofline  top level = the WHOLE STEP against the measured HBM peak in MEASURED_PEAKS.json:
            512 algorithmic bytes per grid-point-step (16A per stage, A = 16*nkx*ny)
            x grid-point-steps/s / n_gpus; `traffic` = real DRAM bytes per step (ncu, profiles/traffic.json).
            `dominant_kernel` = the kernel with the largest share of the step: its contract bytes per
            launch (the fused column kernel is charged 10A: fusing (iii)+(i) removes one re-read of
            omega_hat), its real DRAM bytes (ncu) and both fractions, from CUDA events around every
            launch on the launching stream (nsb_profile_begin/_end).
  secondary the same measurement at 4096^2 (BASELINE config 3), device-resident.
  --gpus N  (torchrun, one rank per GPU) the SAME grid slab-decomposed over N GPUs (strong
            scaling): kx-column blocks in spectral space, y-row blocks in physical space, the
            FFT transposes as NVLink peer stores fused into the kernels; ranks exchange CUDA IPC
            handles over the gloo group once, the data path uses no collective library.
  cpu_baseline  oracle/_ref (the unmodified reference sources + shim FFT; FFTW is not
            installable offline) timed on this box's host cores on a bounded 4096^2 sample.

`--impl reference` times that CPU implementation instead (rank 0 only).
Initial condition on both arms: SURVEY 8d random-phase spectrum (splitmix64 stream, seed 7, k0 = 8,
max|omega| = 5) -- ns_spectral_initialize_random_phase on the GPU arm, the oracle build's
nsref_random_phase_init on the CPU arm (same definition, tests/test_host_api.py pins them together).
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
INIT = dict(seed=7, k0=8.0, max_omega=5.0)
RE = 1.0e4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def all_checksums(dist, xor, add):
    """XOR / wrapping-sum combine of the per-rank state fingerprints (nsb_state_checksum)."""
    if dist is None:
        return xor, add
    import torch
    mine = torch.tensor([xor >> 32, xor & 0xFFFFFFFF, add >> 32, add & 0xFFFFFFFF], dtype=torch.int64)
    parts = [torch.zeros(4, dtype=torch.int64) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    x, a = 0, 0
    for p in parts:
        v = [int(t) for t in p.tolist()]
        x ^= (v[0] << 32) | v[1]
        a = (a + ((v[2] << 32) | v[3])) % 2**64
    return x, a


# ----------------------------------------------------------------------------- CPU arm

def cpu_reference_run(n, steps, warmup):
    """oracle/_ref (kind 'reference') or the C port, all host threads, forced workload. The only place
    bench.py touches oracle/ (the contract's cpu_baseline / reference-arm leg)."""
    import orc
    kind = "ref" if orc.have("ref") else "port"
    if not orc.have(kind):
        orc.build_oracles()
    cores = orc.set_threads(os.cpu_count() or 1)
    o = orc.Oracle(kind, n, n, Re=RE, flavour="timing")
    o.set_hyperviscosity(1, HV["coefficient"], HV["order"], HV["cutoff_ratio"])
    o.random_phase(seed=INIT["seed"], k0=INIT["k0"], max_omega=INIT["max_omega"])
    o.init_forcing("white", kmin=FORCING["kmin"], kmax=FORCING["kmax"], energy_target=FORCING["energy_target"],
                   tau=FORCING["tau"], seed=FORCING["seed"])
    for _ in range(warmup):
        o.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step()
    dt = time.perf_counter() - t0
    build = orc.build_flavour(kind)
    o.close()
    label = ("reference solver (unmodified /root/reference sources) + shim FFT (FFTW unavailable offline; "
             "/root/reference is not on the GPU box, so the library is the one prebuilt in the build container, "
             f"{build})" if kind == "ref" else "C port of the reference + own FFT")
    return {"value": n * n * steps / dt, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port",
            "sample": f"{n}x{n} forced turbulence (same initial condition, forcing and hyperviscosity as the GPU "
                      f"arm), {steps} steps after {warmup} warm-up; {label}",
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
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"forced 2-D turbulence {args.n}x{args.n} FP64 (config 4); timed on a bounded "
                                   f"{n}x{n} sample of it (config 3's grid)", "Re": RE, "hyperviscosity": HV,
                       "forcing": "white", "init": "random-phase spectrum (SURVEY 8d), seed 7"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

class GpuCase:
    """One forced-turbulence solver on this rank's GPU (slab-decomposed when world > 1), driven through
    the reference-facing C entry points of lib/libns_spectral.so."""

    def __init__(self, host, nsblib, n, rank, world, dist, device, forced=True):
        import hostlib
        import ns_structs as st
        self.host, self.nsblib, self.n, self.world = host, nsblib, n, world
        s = hostlib.Solver(host, n, n, Re=RE)
        self.s = s
        host.ns_spectral_backend_context.restype = C.c_void_p
        host.ns_spectral_backend_context.argtypes = [C.c_void_p]
        self.ctx = C.c_void_p(host.ns_spectral_backend_context(s.sp))
        if world > 1:
            # slab decomposition: partition, then swap the CUDA IPC handles of T, T' and the scalar block
            import torch
            nsblib.nsb_mgpu_setup.argtypes = [C.c_void_p, C.c_int, C.c_int]
            nsblib.nsb_mgpu_export.argtypes = [C.c_void_p, C.c_char_p]
            nsblib.nsb_mgpu_import.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
            self._ck(nsblib.nsb_mgpu_setup(self.ctx, rank, world))
            buf = C.create_string_buffer(192)
            self._ck(nsblib.nsb_mgpu_export(self.ctx, buf))
            mine = torch.tensor(list(buf.raw), dtype=torch.uint8)
            gathered = [torch.zeros(192, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(gathered, mine)
            for peer in range(world):
                if peer != rank:
                    self._ck(nsblib.nsb_mgpu_import(self.ctx, peer, bytes(gathered[peer].tolist())))
            dist.barrier()
        host.ns_spectral_set_sync_policy(s.sp, 1)          # LAZY: omega_hat lives in HBM from here on
        assert host.ns_spectral_initialize_random_phase(s.sp, INIT["seed"], INIT["k0"], INIT["max_omega"]) == 0
        h = st.Hyperviscosity()
        host.ns_spectral_hyperviscosity_default(C.byref(h))
        h.enabled, h.coefficient, h.order, h.cutoff_ratio = 1, HV["coefficient"], HV["order"], HV["cutoff_ratio"]
        host.ns_spectral_hyperviscosity_set(s.sp, C.byref(h))
        self.fs, self.force_bytes = None, 0
        if forced:
            cfg = st.ForcingConfig()
            host.ns_forcing_config_default(C.byref(cfg))
            cfg.kmin, cfg.kmax, cfg.energy_target = FORCING["kmin"], FORCING["kmax"], FORCING["energy_target"]
            cfg.correlation_time, cfg.seed, cfg.type = FORCING["tau"], FORCING["seed"], 0  # every rank replays the same stream
            self.fs = st.ForcingState()
            assert host.ns_forcing_state_init(C.byref(self.fs), C.byref(cfg), s.s.k2, s.s.nkx, s.s.nky) == 0
            ib, jb = C.c_int(), C.c_int()
            host.ns_forcing_box_extent(C.byref(self.fs), C.byref(ib), C.byref(jb))
            self.force_bytes = (2 * jb.value + 1) * (ib.value + 1) * 16
        self.steps_taken = 0

    def _ck(self, rc):
        if rc != 0:
            raise SystemExit("bench.py: " + self.nsblib.nsb_last_error().decode())

    def run(self, k):
        assert self.host.ns_spectral_run_steps(self.s.sp, k, C.byref(self.fs) if self.fs is not None else None) == 0, \
            self.host.ns_spectral_last_error()
        self.steps_taken += k

    def timed(self, k, dist):
        """exactly k steps, CUDA events on the solver's stream, barrier + sync on both sides"""
        if dist is not None:
            dist.barrier()
        self.nsblib.nsb_sync(self.ctx)
        l0 = self.nsblib.nsb_launch_count(self.ctx)
        self.nsblib.nsb_timer_start(self.ctx)
        self.run(k)
        ms = C.c_float()
        self.nsblib.nsb_timer_stop(self.ctx, C.byref(ms))
        self.nsblib.nsb_sync(self.ctx)
        if dist is not None:
            dist.barrier()
        return all_max(dist, ms.value / k), self.nsblib.nsb_launch_count(self.ctx) - l0

    def per_kernel(self, steps=2):
        self.nsblib.nsb_profile_begin(self.ctx)
        self.run(steps)
        kms, kcnt = (C.c_double * 6)(), (C.c_int * 6)()
        self.nsblib.nsb_profile_end(self.ctx, kms, kcnt)
        return [kms[k] for k in range(6)], [kcnt[k] for k in range(6)], steps

    def checksum(self):
        out = (C.c_ulonglong * 2)()
        self.nsblib.nsb_state_checksum.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        self._ck(self.nsblib.nsb_state_checksum(self.ctx, out))
        return int(out[0]), int(out[1])

    def plan(self):
        plan = (C.c_int * 8)()
        self.nsblib.nsb_get_plan(self.ctx, plan)
        return {"Y1": plan[0], "Y2": plan[1], "X1": plan[2], "X2": plan[3]}

    def close(self):
        if self.fs is not None:
            self.host.ns_forcing_state_free(C.byref(self.fs))
        self.s.close()


def step_roofline(n, ms_step, world, peak):
    gps = float(n) * n / (ms_step * 1e-3)
    ach = 512.0 * gps / 1e9 / world
    return ach, ach / peak


def load_traffic(n):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tpath)).get(str(n), {})
    except Exception:
        return {}


def run_gpu_arm(args):
    import __graft_entry__ as ge
    import hostlib
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
    peak, peak_src = measured_peaks()

    forced = args.workload == "forced"
    g = GpuCase(host, nsblib, n, rank, world, dist, device, forced=forced)
    g.run(args.warmup)

    # ---- device-resident timed region: exactly K steps
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    ms_step, launches = g.timed(args.steps, dist)
    clocks = sampler.stop() if rank == 0 else None
    total_points = float(n) * n                       # ONE grid, slab-decomposed over the ranks
    value = total_points / (ms_step * 1e-3)

    # ---- e2e: one step per call through the C entry points, forcing H2D + scalars D2H each step
    e2e_steps = max(1, min(args.steps, 10))
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g.run(1)
    nsblib.nsb_sync(g.ctx)
    e2e_s = all_max(dist, (time.perf_counter() - t0) / e2e_steps)
    e2e = {"value": total_points / e2e_s, "unit": UNIT, "h2d_bytes_per_step": g.force_bytes,
           "d2h_bytes_per_step": 2112, "steps": e2e_steps,
           "api": "ns_spectral_run_steps(data, 1, forcing) per step, NS_SYNC_LAZY"}

    # ---- per-kernel timing for the roofline (separate short pass, CUDA events per launch)
    kms, kcnt, prof_steps = g.per_kernel(2)
    names = ["k_alpha1024<inv> (stage-1 spectral ops + inverse y FFT)", "k_chain_B_tma", "k_chain_G",
             "k_chain_F", "k_alpha1024<fwd,inv> (forward y FFT + RK update + next inverse y FFT)", "other"]
    if n <= 1024:
        names = ["k_alpha1024<inv>", "k_beta_inv (unused)", "k_gamma1024", "k_beta_fwd (unused)",
                 "k_alpha1024<fwd,inv>", "other"]
    A = 16.0 * (n // 2 + 1) * n
    # SURVEY 8d contract: 16A per stage = 64A per step (full, unpruned arrays): (i) spectral ops + inverse y
    # pass 1A in / 4A out; (ii) inverse x + Jacobian + forward x 4A in / 1A out; (iii) forward y pass + RK 4A in
    # / 2A out. This decomposition: stage-1 alpha_inv = (i) = 5A; fused alpha = 3 x [(iii)+(i) minus the re-read
    # of omega_hat that the fusion removes = 10A] + 1 x (iii) = 36A; G = 4 x (ii) = 20A. B and F are passes the
    # model does not grant: they are charged NOTHING (contract fraction 0), only their real traffic is shown.
    alg_step = [5 * A, 0.0, 20 * A, 0.0, 36 * A]
    traffic = load_traffic(n) if world == 1 else {}
    per_kernel = {}
    for k, nm in enumerate(names):
        lp = int(kcnt[k]) // prof_steps
        mpl = (kms[k] / kcnt[k]) if kcnt[k] else None
        ent = {"launches_per_step": lp, "ms_per_launch": mpl, "ms_per_step": kms[k] / prof_steps}
        if k < 5 and lp and mpl:
            ent["contract_bytes_per_launch"] = alg_step[k] / lp / world
            ent["contract_frac"] = ent["contract_bytes_per_launch"] / (mpl * 1e-3) / 1e9 / peak
            real = traffic.get(str(k))
            if real:
                ent["real_dram_bytes_per_launch"] = real
                ent["real_frac"] = real / (mpl * 1e-3) / 1e9 / peak
        per_kernel[nm] = ent
    top_i = max(range(5), key=lambda k: kms[k])
    top = dict(per_kernel[names[top_i]])
    top["name"] = names[top_i]
    top["share_of_step"] = kms[top_i] / max(sum(kms[:5]), 1e-30)
    ach, frac = step_roofline(n, ms_step, world, peak)
    real_step = None
    if traffic:
        try:
            real_step = sum(float(traffic[str(k)]) * (int(kcnt[k]) // prof_steps) for k in range(5) if str(k) in traffic)
        except Exception:
            real_step = None
    roofline = {"bound": "hbm", "scope": "whole RK4 step (the dominant kernel alone is in dominant_kernel)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": frac,
                "traffic": real_step, "traffic_note": "real DRAM bytes per step = sum over kernels of ncu dram__bytes "
                "(read+write) per launch x launches per step (profiles/traffic.json); contract = 512 B x nx x ny",
                "algorithmic_bytes_per_step": 512.0 * total_points / world,
                "peak_source": peak_src, "frac_of_8TBs_nominal": ach / 8000.0,
                "dominant_kernel": top, "per_kernel": per_kernel,
                "per_kernel_note": "CUDA events around every launch need one stream: measured in a separate pass "
                                   "with the chain in ONE slab range; in the timed region B, G, F may run as several "
                                   "slab ranges on two streams (profiles/README.md has the ncu launch list)"}
    if world > 1:
        # NVLink: every stage pushes 4 T fields + 1 T' field of the retained columns; (world-1)/world of
        # a rank's share leaves the GPU. SURVEY 8e contract figure (unpruned) alongside the real bytes.
        ncol = (2 * (n // 2 + 1)) // 3 + 1
        real = 4 * 5 * (16.0 * ncol * n / world) * (world - 1) / world
        contract = 20 * (A / world) * (world - 1) / world
        nv = {"bytes_sent_per_gpu_per_step": real, "contract_bytes_per_gpu_per_step": contract,
              "achieved_GBs": real / (ms_step * 1e-3) / 1e9,
              "contract_GBs": contract / (ms_step * 1e-3) / 1e9,
              "peak_GBs": 770.0, "peak_source": "measured peer copy per direction (B200_PROFILING.md)",
              "frac": real / (ms_step * 1e-3) / 1e9 / 770.0,
              "frac_by_contract_bytes": contract / (ms_step * 1e-3) / 1e9 / 770.0}
        # per kernel: the column kernels push the 4 T fields (80 % of the bytes), F pushes T'
        t_bytes, tp_bytes = real * 0.8 / 4.0, real * 0.2 / 4.0
        a_ms = [per_kernel[names[i]]["ms_per_launch"] for i in (0, 4) if per_kernel[names[i]]["ms_per_launch"]]
        f_ms = per_kernel[names[3]]["ms_per_launch"]
        if a_ms:
            nv["column_kernel_push_GBs"] = t_bytes / (min(a_ms) * 1e-3) / 1e9
        if f_ms:
            nv["F_push_GBs"] = tp_bytes / (f_ms * 1e-3) / 1e9
        roofline["nvlink"] = nv

    # fingerprint of the whole run: t and dt after all steps depend on max|u| over the whole field at every
    # step, and the XOR / wrapping sum of the bit patterns of omega_hat over all ranks' own columns pins the
    # FIELD: runs with the same --warmup/--steps must agree bit for bit across GPU counts
    cx, cs = g.checksum()
    cx, cs = all_checksums(dist, cx, cs)
    fingerprint = {"t": float(g.s.s.t).hex(), "dt": float(g.s.s.dt).hex(), "t_dec": g.s.s.t, "dt_dec": g.s.s.dt,
                   "omega_hat_xor": "%016x" % cx, "omega_hat_bitsum": "%016x" % cs,
                   "steps_taken": g.steps_taken}

    # ---- strict drop-in loop (EAGER): bounded, single GPU only, small step count
    e2e_eager = None
    if world == 1 and args.eager_steps > 0:
        host.ns_spectral_set_sync_policy(g.s.sp, 0)
        t0 = time.perf_counter()
        for _ in range(args.eager_steps):
            g.s.s.dt = host.ns_spectral_compute_cfl_timestep(g.s.sp)
            host.ns_spectral_advance_rk4(g.s.sp)
        te = (time.perf_counter() - t0) / args.eager_steps
        m_bytes, n_bytes = 16 * (n // 2 + 1) * n, 8 * n * n
        e2e_eager = {"value": float(n) * n / te, "unit": UNIT, "steps": args.eager_steps,
                     "h2d_bytes_per_step": m_bytes + 2 * n_bytes, "d2h_bytes_per_step": 4 * m_bytes + 3 * n_bytes,
                     "api": "ns_spectral_compute_cfl_timestep + ns_spectral_advance_rk4, NS_SYNC_EAGER (pageable host arrays)"}
    plan = g.plan()
    g.close()

    # ---- secondary: BASELINE config 3's grid (4096^2), same workload, device resident
    secondary = None
    if args.secondary_n > 0 and args.secondary_n != n:
        m = args.secondary_n
        g2 = GpuCase(host, nsblib, m, rank, world, dist, device)
        g2.run(max(3, args.warmup))
        k2 = max(args.steps, 50)
        ms2, _ = g2.timed(k2, dist)
        a2, f2 = step_roofline(m, ms2, world, peak)
        c2x, c2s = g2.checksum()
        c2x, c2s = all_checksums(dist, c2x, c2s)
        secondary = {"workload": f"forced 2-D turbulence {m}x{m} FP64 (BASELINE config 3), same forcing / "
                                 "hyperviscosity / initial condition", "value": float(m) * m / (ms2 * 1e-3),
                     "unit": UNIT, "ms_per_step": ms2, "steps": k2, "plan": g2.plan(),
                     "roofline_step": {"achieved": a2, "frac": f2, "peak": peak, "unit": "GB/s"},
                     "fingerprint": {"t": float(g2.s.s.t).hex(), "dt": float(g2.s.s.dt).hex(),
                                     "omega_hat_xor": "%016x" % c2x, "omega_hat_bitsum": "%016x" % c2s}}
        g2.close()

    if rank != 0:
        return
    cpu = None
    if args.cpu_steps > 0:
        r = cpu_reference_run(args.cpu_n, args.cpu_steps, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"forced 2-D turbulence {n}x{n} FP64, hyperviscosity + white forcing, adaptive CFL "
                                    f"dt (BASELINE config {'4' if n == 16384 else '3' if n == 4096 else '5 grid, forced' if n == 32768 else 'n=' + str(n)})")
                       if forced else
                       (f"decaying 2-D turbulence {n}x{n} FP64, hyperviscosity, no forcing, adaptive CFL dt (BASELINE config "
                        f"{'5' if n == 32768 else '2' if n == 1024 else 'n=' + str(n)})"),
                       "Re": RE, "hyperviscosity": HV, "forcing": {"type": "white", **FORCING} if forced else {"type": "none"},
                       "init": "ns_spectral_initialize_random_phase: spectrum k^2 exp(-(k/8)^2), splitmix64 phases "
                               "(seed 7), max|omega| = 5 (SURVEY 8d)",
                       "parallelism": "1 GPU" if world == 1 else
                       f"slab decomposition over {world} GPUs: kx-column blocks / y-row blocks, FFT transposes as "
                       "NVLink peer stores fused into the kernels (one process per GPU, CUDA IPC)",
                       "plan": plan,
                       "l2": "inputs larger than L2: working set >> 126 MB at this size, no flush needed"
                             if n >= 2048 else "working set fits L2 at this size"},
            "clocks": clocks, "e2e": e2e, "e2e_eager": e2e_eager, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary, "fingerprint": fingerprint}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", "--grid", dest="n", type=int, default=16384,
                    help="grid size (nx = ny); under torchrun use --grid (its parser takes --n for its own options)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-n", type=int, default=4096, help="grid of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=3, help="timed CPU steps (0 = skip cpu_baseline)")
    ap.add_argument("--eager-steps", type=int, default=1, help="strict EAGER drop-in steps (0 = skip)")
    ap.add_argument("--secondary-n", type=int, default=4096, help="grid of the secondary result (0 = skip)")
    ap.add_argument("--workload", default="forced", choices=["forced", "decaying"],
                    help="forced: BASELINE configs 3/4 (white forcing); decaying: configs 2/5 (no forcing)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
