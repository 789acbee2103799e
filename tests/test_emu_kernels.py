"""Kernel-logic tests that run WITHOUT a GPU: the product's .cu sources compiled for the
CPU emulator (tests/emu) checked against the CPU oracle at small sizes, including
factorisations that mimic the 16384^2 plan (Y2 = X1 = 16) in miniature."""
import pytest

import parity_cases as pc
from conftest import best_oracle_kind

HV = dict(enabled=1, coefficient=0.5, order=4, cutoff_ratio=0.3)
SFD = dict(enabled=1, coefficient=0.5, order=4, cutoff_ratio=0.3, sfd_enabled=1, sfd_strength=0.3,
           sfd_timescale=0.4)

PLANS = [
    (32, 32, {}),
    (64, 32, dict(y1=8, x2=8)),
    (64, 64, dict(y1=16, x2=32)),
    (128, 64, dict(y1=4, x2=8, col_tile=4, beta_batch=4)),
    (64, 128, dict(y1=128, x2=4)),
]


@pytest.mark.parametrize("nx,ny,plan", PLANS)
def test_emu_transforms(nsb, emu_lib, oracles, nx, ny, plan):
    pc.check_transforms(nsb, emu_lib, oracles, best_oracle_kind(oracles), nx, ny, **plan)


@pytest.mark.parametrize("nx,ny,plan", PLANS[:3])
def test_emu_operators(nsb, emu_lib, oracles, nx, ny, plan):
    pc.check_operators(nsb, emu_lib, oracles, best_oracle_kind(oracles), nx, ny, **plan)


def test_emu_step_plain(nsb, emu_lib, oracles):
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), 32, 32, 3)


def test_emu_step_taylor_green_host_dt(nsb, emu_lib, oracles):
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), 32, 32, 2, init="tg", Re=100.0,
                      device_dt=False)


def test_emu_step_forced_white_split(nsb, emu_lib, oracles):
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), 64, 32, 2, hv=HV,
                      forcing=dict(kind="white"), y1=8, x2=16)


def test_emu_step_forced_colored_sfd_split(nsb, emu_lib, oracles):
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), 32, 64, 2, hv=SFD,
                      forcing=dict(kind="colored"), y1=16, x2=8, col_tile=4, beta_batch=4)


def test_emu_free_running(nsb, emu_lib, oracles):
    pc.check_free_running(nsb, emu_lib, oracles, best_oracle_kind(oracles), 32, 32, 20, every=5, hv=HV)


@pytest.mark.parametrize("nx,ny,plan", [(1024, 16, {}), (2048, 16, dict(x2=1024))])
def test_emu_gamma1024_fast_path(nsb, emu_lib, oracles, nx, ny, plan):
    """X2 == 1024 selects the warp-level register FFT kernel (k_gamma1024)."""
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), nx, ny, 2, hv=HV,
                      forcing=dict(kind="white"), **plan)


@pytest.mark.parametrize("nx,ny,plan,hv,forcing", [
    (16, 1024, {}, HV, "white"),
    (32, 2048, dict(y1=1024), SFD, "colored"),
])
def test_emu_alpha1024_fused_path(nsb, emu_lib, oracles, nx, ny, plan, hv, forcing):
    """Y1 == 1024 selects the fused column kernel (k_alpha1024) and the K-layout state."""
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), nx, ny, 2, hv=hv,
                      forcing=dict(kind=forcing), **plan)


def test_emu_alpha1024_free_running_with_diagnostics(nsb, emu_lib, oracles):
    pc.check_free_running(nsb, emu_lib, oracles, best_oracle_kind(oracles), 16, 1024, 6, every=3, hv=HV)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("nx,ny,plan", [
    (2048, 32, dict(x2=1024, y1=16)),          # X1 = 2, Y2 = 2: 8 y1 values per slab, 2 slabs
    (4096, 16, dict(x2=1024, y1=4)),           # X1 = 4, Y2 = 4: one slab
    (1024, 32, dict(x2=1024, y1=2)),           # X1 = 1, Y2 = 16: comb pairs are row pairs
])
def test_emu_chain(nsb, emu_lib, oracles, monkeypatch, mode, nx, ny, plan):
    """X2 == 1024 with a beta split selects the B/G/F chain (nsb_chain.cuh): mode 1 = three
    plain kernels, mode 2 = the persistent dataflow kernel over ring buffers."""
    monkeypatch.setenv("NSB_CHAIN", str(mode))
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), nx, ny, 2, hv=HV,
                      forcing=dict(kind="white"), **plan)


@pytest.mark.parametrize("nx,ny,plan", [
    (2048, 64, dict(x2=1024, y1=2)),           # X1 = 2, Y2 = 32: 32-row slabs, one y1 value each, 2 slabs
    (32768, 16, dict(x2=1024, y1=4)),          # X1 = 32, Y2 = 4: 22 column segments, 4-column B batches
])
def test_emu_chain_32_point_factors(nsb, emu_lib, oracles, monkeypatch, nx, ny, plan):
    """32768 = 1024 x 32: the chain with 32-point short factors (BASELINE config 5's factorisation),
    plus the generic c2r / r2c kernels at the same factorisation (k_beta_*<32, .> / <., 32>).
    F runs with the 8-column batches it takes at 32768^2 (automatic only when its tile exceeds 100 KB)."""
    monkeypatch.setenv("NSB_CHAIN_FBC", "8")
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), nx, ny, 1 if nx > 4096 else 2, hv=HV,
                      forcing=dict(kind="white"), **plan)
    pc.check_transforms(nsb, emu_lib, oracles, best_oracle_kind(oracles), nx, ny, **plan)


def test_emu_chain_ring_reuse(nsb, emu_lib, oracles, monkeypatch):
    """4 slabs through 2-slot rings: slot reuse and the completion-counter waits."""
    monkeypatch.setenv("NSB_CHAIN", "2")
    monkeypatch.setenv("NSB_CHAIN_SLOTS_Z", "2")
    monkeypatch.setenv("NSB_CHAIN_SLOTS_P", "2")
    monkeypatch.setenv("NSB_CHAIN_LEAD_G", "1")
    monkeypatch.setenv("NSB_CHAIN_LEAD_F", "1")
    pc.check_per_step(nsb, emu_lib, oracles, best_oracle_kind(oracles), 2048, 64, 2, hv=HV,
                      forcing=dict(kind="white"), x2=1024, y1=32)


def test_emu_chain_matches_plain_kernels(nsb, emu_lib, monkeypatch):
    """Plain and persistent chain against each other and against the separate
    beta/gamma kernels: only the Hermitian packing differs (psi_x + i omega_y instead of
    psi_x + i psi_y), i.e. rounding-level differences."""
    import numpy as np
    rng = np.random.default_rng(5)
    nx, ny = 2048, 32
    w = np.zeros((ny, nx // 2 + 1), dtype=np.complex128)
    w[:, :40] = 0.01 * (rng.standard_normal((ny, 40)) + 1j * rng.standard_normal((ny, 40)))
    w[0, 0] = 0
    out = {}
    for mode in (0, 1, 2):
        monkeypatch.setenv("NSB_CHAIN", str(mode))
        c = nsb.NsbContext(nx, ny, Re=500.0, lib=emu_lib, x2=1024, y1=16)
        c.upload(nsb.OMEGA_HAT, w)
        c.apply_dealiasing(nsb.OMEGA_HAT)
        c.step(0.0)
        out[mode] = (c.download(nsb.OMEGA_HAT), c.state()[0])
        c.close()
    # plain (packing in front of the k1 DFT since round 2) vs persistent (behind it): rounding level
    assert pc.rel(out[1][0], out[2][0]) < 1e-14 and abs(out[1][1] - out[2][1]) <= 1e-15 * out[2][1]
    assert pc.rel(out[1][0], out[0][0]) < 1e-14
    assert abs(out[1][1] - out[0][1]) <= 1e-15 * out[0][1]


def test_emu_slab_decomposition_two_devices(nsb, emu_lib):
    """Two slab-decomposed contexts (kx-column blocks / y-row blocks, transposes as pushes into
    the owner's buffers) against one context: the arithmetic per mode is the same, so
    omega_hat and the device-side CFL dt must be bit-identical."""
    import numpy as np
    nx, ny = 1024, 2048
    rng = np.random.default_rng(9)
    w = np.zeros((ny, nx // 2 + 1), dtype=np.complex128)
    kk = 24
    blk = rng.standard_normal((2 * kk + 1, kk + 1)) + 1j * rng.standard_normal((2 * kk + 1, kk + 1))
    w[(np.arange(-kk, kk + 1) % ny)[:, None], np.arange(kk + 1)[None, :]] = 0.02 * blk
    w[0, 0] = 0.0
    col = w[:, 0].copy()
    w[:, 0] = 0.5 * (col + np.conj(col[(-np.arange(ny)) % ny]))
    one = nsb.NsbContext(nx, ny, Re=500.0, lib=emu_lib)
    one.upload(nsb.OMEGA_HAT, w)
    one.set_hyperviscosity(**HV)
    one.step(0.0)
    ref, dt_ref = one.download(nsb.OMEGA_HAT), one.state()[0]
    ck_ref = one.state_checksum()
    one.close()
    grp = nsb.NsbGroup(nx, ny, 2, devices=[0, 0], lib=emu_lib, Re=500.0)
    grp.upload(nsb.OMEGA_HAT, w)
    grp.each(lambda c: c.set_hyperviscosity(**HV))
    parts = grp.each(lambda c: c.partition())
    assert parts[0]["col1"] == parts[1]["col0"] and parts[1]["col1"] == 343 and parts[1]["row0"] == 1024
    grp.step(0.0)
    got = grp.download_omega_hat()
    dts = grp.each(lambda c: c.state()[0])
    sums = grp.each(lambda c: c.state_checksum())
    grp.close()
    assert np.array_equal(got, ref)
    assert dts[0] == dts[1] == dt_ref
    # nsb_state_checksum: XOR / wrapping sum over the own columns combine to the one-device value
    assert sums[0][0] ^ sums[1][0] == ck_ref[0] and (sums[0][1] + sums[1][1]) % 2**64 == ck_ref[1]
    bits = ref[:, :343].copy().view(np.uint64).reshape(ny, 343, 2)
    re, im = bits[..., 0], bits[..., 1]
    rot = (im << np.uint64(1)) | (im >> np.uint64(63))
    assert ck_ref[0] == int(np.bitwise_xor.reduce((re ^ rot).reshape(-1)))
    assert ck_ref[1] == (int(re.sum(dtype=np.uint64)) + 3 * int(im.sum(dtype=np.uint64))) % 2**64


def test_emu_spectral_stats(nsb, emu_lib):
    pc.check_spectral_stats(nsb, emu_lib, 64, 32)


def test_emu_parseval_diagnostics(nsb, emu_lib):
    pc.check_parseval(nsb, emu_lib, 64, 32)
    pc.check_parseval(nsb, emu_lib, 32, 64, y1=8, x2=8)
