import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import spectra_b200 as sb

        sb.device_info()
        return True
    except Exception:
        return False


@pytest.fixture(scope="session", autouse=True)
def _build_cpu_side():
    """The oracle (test infrastructure) and the synthetic generator are plain C/C++: build on demand."""
    import oracle
    from spectra_b200 import synth

    oracle.build()
    synth.build()


@pytest.fixture(scope="session")
def gpu():
    if not _has_gpu():
        pytest.fail("a test marked `gpu` ran without a usable CUDA device / built library (no CPU fallback exists)")
    import spectra_b200 as sb

    return sb
