import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_pieces_built():
    """The C-ABI library, the CLI, the oracle and the generator are build products (git-ignored): build them once
    per session when any is missing (nvcc cross-compiles without a GPU)."""
    from nomalise_kmers_multi_large_b200 import capi
    needed = [capi.LIB_PATH, capi.CLI_PATH, ROOT / "oracle" / "libnk_oracle.so", ROOT / "oracle" / "nk_oracle",
              ROOT / "tools" / "libnk_synth.so"]
    if not all(p.exists() for p in needed):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def emu_lib():
    """Test-only CPU emulation of the device engine (tests/emu), same C ABI as the CUDA library."""
    import ctypes
    from nomalise_kmers_multi_large_b200 import capi
    subprocess.run(["make", "-C", str(ROOT / "tests" / "emu")], check=True, capture_output=True)
    lib = ctypes.CDLL(str(ROOT / "tests" / "emu" / "libnk_emu.so"))
    return capi._declare_engine(lib)


@pytest.fixture(scope="session")
def cuda_lib():
    from nomalise_kmers_multi_large_b200 import capi
    return capi.load_library()
