"""Drop-in boundary checks that need no GPU: every symbol include/*.h declares is exported
by the built libraries, the host-side forcing generator replays the reference's seeded
stream bit-exactly, ns_config behaves like the reference's on its own fixture files, and
the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import hostlib
import orc
from conftest import ROOT, _clean_env

GOLD = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "navier-stokes-solvers_b200")


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-s", "-C", PKG, "all"], check=True, env=_clean_env(), stdout=subprocess.DEVNULL)
    return True


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(built):
    nsb = C.CDLL(os.path.join(PKG, "lib", "libnsb.so"))
    for name in _declared("nsb.h", "nsb_"):
        assert hasattr(nsb, name), name
    host = C.CDLL(hostlib.PRODUCT)
    for hdr, pfx in [("ns_spectral_solver.h", "ns_spectral_"), ("ns_forcing.h", "ns_forcing_"),
                     ("ns_config.h", "ns_config_"), ("ns_statistics.h", "ns_s"), ("ns_vtk_writer.h", "ns_")]:
        names = _declared(hdr, pfx)
        assert len(names) >= 2
        for name in names:
            assert hasattr(host, name), name
    assert set(hostlib.NSB_SYMBOLS) == set(_declared("nsb.h", "nsb_"))
    # the 22 reference prototypes (ns_spectral_solver.h:71-113) + 4 forcing + 10 config entry points
    assert len(hostlib.SPECTRAL_SYMBOLS[:24]) == 24 and set(hostlib.SPECTRAL_SYMBOLS) <= set(
        _declared("ns_spectral_solver.h", "ns_spectral_"))


def test_struct_layout_is_reference_layout_plus_backend():
    assert hostlib.ProductStruct.backend.offset == C.sizeof(orc.SpectralStruct)
    for name, _ in orc._SPECTRAL_FIELDS:
        assert getattr(hostlib.ProductStruct, name).offset == getattr(orc.SpectralStruct, name).offset


def test_no_cpu_fallback_without_gpu(built, nsb):
    lib = nsb.load_nsb()
    if lib.nsb_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(nsb.NsbError, match="no CUDA device"):
        nsb.NsbContext(32, 32, lib=lib)
    host = hostlib.load(hostlib.PRODUCT)
    sp = host.ns_spectral_allocate(32, 32, 6.28, 6.28, 100.0)
    host.ns_spectral_initialize_grids(sp)
    host.ns_spectral_initialize_fft_plans(sp)          # reports on stderr, leaves plans NULL
    assert not sp.contents.forward_plan
    assert b"no CUDA device" in host.ns_spectral_last_error()
    host.ns_spectral_free(sp)


def test_missing_library_fails_loudly(nsb, tmp_path):
    with pytest.raises(nsb.NsbError, match="no CPU fallback"):
        nsb.load_nsb(str(tmp_path / "libnsb.so"))


def test_product_forcing_replays_reference_stream_bit_exactly(built):
    host = hostlib.load(hostlib.PRODUCT)
    gold = json.load(open(os.path.join(GOLD, "forcing_kat.json")))
    nkx = nky = 4
    k2 = np.array([[float(i * i + j * j) for i in range(nkx)] for j in range(nky)])
    k2p = k2.ctypes.data_as(hostlib.c_double_p)
    for name, kid, tau in [("white", 0, None), ("colored", 1, 0.2)]:
        for use_box in (False, True):
            cfg = orc.ForcingConfig()
            host.ns_forcing_config_default(C.byref(cfg))
            assert (cfg.kmin, cfg.kmax, cfg.energy_target, cfg.correlation_time, cfg.seed, cfg.type) == \
                   (0.0, 4.0, 1.0, 0.5, 0, 0)                      # forcing.c:41-51
            cfg.kmin, cfg.kmax, cfg.energy_target, cfg.seed, cfg.type = 0.5, 2.5, 1.5, 42, kid
            if tau:
                cfg.correlation_time = tau
            st = orc.ForcingState()
            assert host.ns_forcing_state_init(C.byref(st), C.byref(cfg), k2p, nkx, nky) == 0
            assert st.active_modes == 7
            idx = (C.c_int * st.active_modes)()
            ib, jb = C.c_int(), C.c_int()
            assert host.ns_forcing_active_indices(C.byref(st), idx) == 0
            assert host.ns_forcing_box_extent(C.byref(st), C.byref(ib), C.byref(jb)) == 0
            for c in gold[name]["calls"]:
                exp = np.array([float.fromhex(h) for h in c["values"]]).view(np.complex128).reshape(4, 4)
                if use_box:
                    # the 4x4 toy grid treats every row as a positive band
                    box = np.zeros((2 * jb.value + 1, ib.value + 1), dtype=np.complex128)
                    host.ns_forcing_generate_box(C.byref(st), c["dt"], idx, ib.value, jb.value,
                                                 box.ctypes.data_as(hostlib.c_double_p))
                    got = np.zeros((4, 4), dtype=np.complex128)
                    for j in range(4):
                        band = j if j <= nky // 2 else j - nky
                        got[j, : ib.value + 1] = box[band + jb.value]
                else:
                    got = np.zeros((4, 4), dtype=np.complex128)
                    host.ns_forcing_generate(C.byref(st), c["dt"], k2p, got.ctypes.data_as(hostlib.c_double_p))
                assert np.array_equal(got, exp), (name, use_box)
                assert int(st.rng_state) == c["rng_state"] and int(st.have_spare) == c["have_spare"]
            host.ns_forcing_state_free(C.byref(st))
    # no active mode -> -1 (forcing.c:102)
    cfg.kmin, cfg.kmax = 100.0, 200.0
    st = orc.ForcingState()
    assert host.ns_forcing_state_init(C.byref(st), C.byref(cfg), k2p, nkx, nky) == -1
    host.ns_forcing_state_free(C.byref(st))


INI = """[solver]
type = spectral

[grid]
nx = 128
ny = 128

[domain]
Lx = 2.0
Ly = 2.0

[fluid]
Re = 750.0

[time]
dt = 0.002
t_end = 1.5

[forcing]
type = white
energy_target = 1.25
correlation_time = 0.4

[spectral.hyperviscosity]
coefficient = 0.015
order = 8
"""
JSON = '{\n  "grid": {\n    "nx": 192,\n    "ny": 96\n  },\n  "solver": {\n    "type": "finite_difference"\n  },\n' \
       '  "time": {\n    "dt": 0.001\n  }\n}\n'


def _schema(entries):
    arr = (hostlib.SchemaEntry * len(entries))()
    keep = []
    for k, (key, typ, req, mn, has_min, allowed) in enumerate(entries):
        arr[k].key, arr[k].type, arr[k].required = key, typ, req
        arr[k].min_value, arr[k].has_min, arr[k].has_max = mn, has_min, 0
        if allowed:
            a = (C.c_char_p * len(allowed))(*allowed)
            keep.append(a)
            arr[k].allowed_values, arr[k].allowed_count = a, len(allowed)
    return arr, keep


def test_config_loader_reference_case(built, tmp_path):
    """tests/test_solvers.c:297-347 with the contents of config/parameters.ini and
    tests/config_override.json (same keys and values)."""
    host = hostlib.load(hostlib.PRODUCT)
    ini = tmp_path / "parameters.ini"
    ini.write_text(INI)
    js = tmp_path / "config_override.json"
    js.write_text(JSON)
    cfg = hostlib.Config()
    host.ns_config_init(C.byref(cfg))
    for k, v in [(b"solver.type", b"spectral"), (b"grid.nx", b"64"), (b"grid.ny", b"64"), (b"time.dt", b"0.005")]:
        assert host.ns_config_set(C.byref(cfg), k, v) == 0
    paths = (C.c_char_p * 2)(str(ini).encode(), str(js).encode())
    err = C.create_string_buffer(256)
    assert host.ns_config_load_files(C.byref(cfg), paths, 2, err, 256) == 0, err.value
    assert host.ns_config_has_key(C.byref(cfg), b"domain.Lx")
    assert host.ns_config_get_string(C.byref(cfg), b"solver.type", b"") == b"finite_difference"
    assert host.ns_config_get_int(C.byref(cfg), b"grid.nx", 0) == 192
    assert host.ns_config_get_int(C.byref(cfg), b"grid.ny", 0) == 96
    assert host.ns_config_get_double(C.byref(cfg), b"domain.Lx", 0.0) == 2.0
    assert host.ns_config_get_double(C.byref(cfg), b"time.dt", 0.0) == 0.001
    assert host.ns_config_get_double(C.byref(cfg), b"forcing.energy_target", 0.0) == 1.25
    assert host.ns_config_get_double(C.byref(cfg), b"spectral.hyperviscosity.coefficient", 0.0) == 0.015
    S, I, D = 0, 1, 2
    schema, keep = _schema([
        (b"solver.type", S, 1, 0.0, 0, [b"spectral", b"finite_difference"]),
        (b"grid.nx", I, 1, 2.0, 1, None), (b"grid.ny", I, 1, 2.0, 1, None),
        (b"domain.Lx", D, 1, 1e-6, 1, None), (b"fluid.Re", D, 1, 1e-6, 1, None),
        (b"time.dt", D, 1, 1e-9, 1, None), (b"time.t_end", D, 1, 1e-9, 1, None),
        (b"forcing.type", S, 0, 0.0, 0, [b"none", b"white", b"colored"]),
        (b"spectral.hyperviscosity.order", I, 0, 0.0, 1, None)])
    assert host.ns_config_validate(C.byref(cfg), schema, len(schema), err, 256) == 0, err.value
    # failures: missing required key, below minimum, value not in the allowed set
    bad, keep2 = _schema([(b"no.such.key", D, 1, 0.0, 0, None)])
    assert host.ns_config_validate(C.byref(cfg), bad, 1, err, 256) == -1
    assert err.value == b"Missing required configuration key: no.such.key"
    bad, keep2 = _schema([(b"grid.nx", I, 1, 1000.0, 1, None)])
    assert host.ns_config_validate(C.byref(cfg), bad, 1, err, 256) == -1
    assert err.value == b"Configuration integer below minimum: grid.nx"
    bad, keep2 = _schema([(b"forcing.type", S, 0, 0.0, 0, [b"none"])])
    assert host.ns_config_validate(C.byref(cfg), bad, 1, err, 256) == -1
    host.ns_config_free(C.byref(cfg))
    assert cfg.count == 0


def test_config_edge_cases(built, tmp_path):
    host = hostlib.load(hostlib.PRODUCT)
    cfg = hostlib.Config()
    host.ns_config_init(C.byref(cfg))
    err = C.create_string_buffer(256)
    ini = tmp_path / "edge.ini"
    ini.write_text("top = 1\n; comment\n# comment\n[ a.b ]\n  key with space =  padded value  \nflag = Yes\nbad = 1x\n")
    js = tmp_path / "edge.JSON"
    js.write_text('{"a": {"b": {"c": 1e-3, "s": "q\\"uo\\nte", "t": true, "n": null}}, "big": 123456789012}')
    paths = (C.c_char_p * 2)(str(ini).encode(), str(js).encode())
    assert host.ns_config_load_files(C.byref(cfg), paths, 2, err, 256) == 0, err.value
    assert host.ns_config_get_int(C.byref(cfg), b"top", 0) == 1
    assert host.ns_config_get_string(C.byref(cfg), b"a.b.key with space", b"") == b"padded value"
    assert host.ns_config_get_bool(C.byref(cfg), b"a.b.flag", 0) == 1
    assert host.ns_config_get_int(C.byref(cfg), b"a.b.bad", -7) == -7          # unparsable -> default
    assert host.ns_config_get_double(C.byref(cfg), b"a.b.c", 0.0) == 1e-3
    assert host.ns_config_get_string(C.byref(cfg), b"a.b.c", b"") == b"0.001"   # numbers re-printed %.15g
    assert host.ns_config_get_string(C.byref(cfg), b"a.b.s", b"") == b'q"uo\nte'
    assert host.ns_config_get_bool(C.byref(cfg), b"a.b.t", 0) == 1
    assert host.ns_config_get_string(C.byref(cfg), b"a.b.n", b"x") == b""       # null -> ""
    assert host.ns_config_get_string(C.byref(cfg), b"big", b"") == b"123456789012"
    assert host.ns_config_get_int(C.byref(cfg), b"missing", 5) == 5
    host.ns_config_free(C.byref(cfg))
    # malformed inputs
    for name, text, msg in [("m1.ini", "[broken\n", b"Malformed section header in INI file"),
                            ("m2.ini", "novalue\n", b"Malformed key-value line in INI file"),
                            ("m3.json", '{"a": [1,2]}', b"Invalid JSON value"),
                            ("m4.json", '{"a": 1} x', b"Unexpected trailing data in JSON file"),
                            ("m5.json", '{"a" 1}', b"Expected ':' in JSON key/value pair")]:
        p = tmp_path / name
        p.write_text(text)
        host.ns_config_init(C.byref(cfg))
        paths1 = (C.c_char_p * 1)(str(p).encode())
        assert host.ns_config_load_files(C.byref(cfg), paths1, 1, err, 256) == -1
        assert err.value.startswith(msg), (name, err.value)
        host.ns_config_free(C.byref(cfg))
    paths1 = (C.c_char_p * 1)(b"/nonexistent/file.ini")
    host.ns_config_init(C.byref(cfg))
    assert host.ns_config_load_files(C.byref(cfg), paths1, 1, err, 256) == -1
    assert err.value.startswith(b"Failed to open INI configuration file")
