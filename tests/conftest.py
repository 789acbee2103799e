import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _clean_env():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


@pytest.fixture(scope="session")
def nsb():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def oracles():
    """Builds oracle/libns_oracle.so (and oracle/_ref when /root/reference exists)."""
    import orc
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True,
                   env=_clean_env(), stdout=subprocess.DEVNULL)
    return orc


@pytest.fixture(scope="session")
def emu_lib(nsb):
    """The kernel sources compiled for the CPU emulator (tests only)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True,
                   env=_clean_env(), stdout=subprocess.DEVNULL)
    return nsb.load_nsb(os.path.join(ROOT, "tests", "emu", "build", "libnsb_emu.so"))


@pytest.fixture(scope="session")
def cuda_lib(nsb):
    """The product library; fails loudly if it is missing or no GPU is visible."""
    lib = nsb.load_nsb()
    assert lib.nsb_device_count() > 0, "no CUDA device visible"
    return lib


def best_oracle_kind(orc):
    return "ref" if orc.have("ref") else "port"
