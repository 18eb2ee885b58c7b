"""pytest configuration: `gpu` marker, repo-root imports, oracle build-on-demand."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing in sphb_create."""
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def graft():
    import __graft_entry__ as g
    return g


@pytest.fixture(scope="session")
def pkg(graft):
    lib = ROOT / "sph-particle-simulator_b200" / "libsphb.so"
    if not lib.exists():
        subprocess.run(["bash", str(ROOT / "sph-particle-simulator_b200" / "csrc" / "build.sh")], check=True)
    return graft.load_package()


@pytest.fixture(scope="session")
def po(graft):
    """The oracle front-end (test infrastructure).  The C restatement is built on demand; the compiled
    reference (oracle/_ref) is used when present and never rebuilt here."""
    if not (ROOT / "oracle" / "liboracle_port.so").exists():
        subprocess.run(["make", "-s", "-f", str(ROOT / "oracle" / "Makefile"), "port"], check=True)
    return graft.load_oracle()


@pytest.fixture(scope="session")
def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
