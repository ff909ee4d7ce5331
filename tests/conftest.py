import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_package():
    """Imports zprize23-gpu-submission_b200/ (hyphenated directory) as module `zprize23_gpu_submission_b200`."""
    name = "zprize23_gpu_submission_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "zprize23-gpu-submission_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def emu_lib(pkg):
    """The product sources compiled against the CPU emulation layer (tests/emu) — unit-test vehicle only."""
    path = pkg._build.build_emu()
    return pkg.load_library(path)


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lib = pkg.load_library()  # fails loudly if the extension is missing
    assert lib.zp_device_available() == 1
    return lib
