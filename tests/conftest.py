import ctypes
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import box2d_b200 as b2  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _ref_path(name: str) -> Path:
	return ROOT / "oracle" / "_ref" / name


@pytest.fixture(scope="session")
def ref_lib():
	"""The untouched reference (CPU solver) built by oracle/build_ref.py -- the parity oracle."""
	path = _ref_path("libbox2d_ref.so")
	if not path.is_file():
		from tools import buildlib
		if not buildlib.reference_available():
			pytest.skip("oracle/_ref not built and /root/reference absent")
		buildlib.build_reference_libs()
	return b2._bind_harness(ctypes.CDLL(str(path)))


@pytest.fixture(scope="session")
def golden_hashes():
	return json.loads((GOLDEN / "hashes.json").read_text())


@pytest.fixture(scope="session")
def capture_files():
	return sorted(GOLDEN.glob("*.b2cap.gz"))


@pytest.fixture(scope="session")
def gpu_host_lib():
	lib = b2.host_lib()
	lib.b2GpuSeam_InstallPinnedAllocator()
	return lib
