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
def emu():
    """The product's kernels compiled for the CPU against the CUDA execution model of tools/cuda_emu (logic checks without a GPU)."""
    import emu_loader

    return emu_loader.load()


@pytest.fixture(params=["forward", "reverse", "rotate"])
def emu_order(request, emu):
    """Runs a test three times: with the fibers of a CTA scheduled in ascending order, in descending order, and ascending from a start
    thread that rotates every scheduling round.  A result that depends on
    the order means code that relies on warp-lockstep execution between two synchronisation points -- a data race under independent
    thread scheduling (this is how the missing __syncwarp() of dense_gen.cu's overflow rescaling was found)."""
    emu.lib().cuda_emu_set_reverse({"forward": 0, "reverse": 1, "rotate": 2}[request.param])
    yield emu
    emu.lib().cuda_emu_set_reverse(0)


@pytest.fixture(scope="session")
def gpu():
    if os.environ.get("SB200_TEST_BACKEND") == "emu":
        # developer mode: run the gpu-marked tests against the kernel-logic emulator (python -m pytest tests -m gpu with this variable set)
        import emu_loader

        return emu_loader.load()
    if not _has_gpu():
        pytest.fail("a test marked `gpu` ran without a usable CUDA device / built library (no CPU fallback exists)")
    import spectra_b200 as sb

    return sb
