"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares,
its Python mirror has the header's layout, and it refuses to run without a device (no CPU fallback)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

import box2d_b200 as b2

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
	text = (ROOT / "include" / "b2_gpu_solver.h").read_text()
	return sorted(set(re.findall(r"B2GPU_API\s+[\w\s\*]+?\b(b2Gpu\w+)\s*\(", text)))


def test_header_declares_entry_points():
	names = _declared_symbols()
	for required in ("b2GpuSolverCreate", "b2GpuSolverDestroy", "b2GpuSolverStep", "b2GpuSolverUpload", "b2GpuSolverRun",
					 "b2GpuSolverDownload", "b2GpuSolverStepBatch", "b2GpuHostAlloc", "b2GpuHostFree"):
		assert required in names


def test_library_exports_every_declared_symbol():
	lib = b2.solver_lib()
	for name in _declared_symbols():
		assert hasattr(lib, name), f"{name} declared in include/b2_gpu_solver.h but not exported"


def test_python_mirror_matches_header_layout(tmp_path):
	src = tmp_path / "sz.c"
	src.write_text(
		'#include "b2_gpu_solver.h"\n#include "b2gpu_layout.h"\n#include <stdio.h>\n#include <stddef.h>\n'
		'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(b2GpuStepDesc), sizeof(b2GpuStepResult),'
		' offsetof(b2GpuStepDesc,states), offsetof(b2GpuStepDesc,colors), offsetof(b2GpuStepDesc,overflow),'
		' sizeof(b2lJointSim));return 0;}\n')
	exe = tmp_path / "sz"
	subprocess.check_call(["gcc", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)])
	got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
	assert got == [ctypes.sizeof(b2.StepDesc), ctypes.sizeof(b2.StepResult), b2.StepDesc.states.offset,
				   b2.StepDesc.colors.offset, b2.StepDesc.overflow.offset, b2.JOINT_SIZE]


def test_no_cpu_fallback_without_device():
	lib = b2.solver_lib()
	if lib.b2GpuGetDeviceCount() > 0:
		pytest.skip("a device is present")
	assert lib.b2GpuSolverCreate(0) is None
	assert b"no CPU fallback" in lib.b2GpuGetLastError()
	with pytest.raises(RuntimeError):
		b2.GpuSolver()


def test_host_library_has_no_cpu_solver_entry():
	"""The product host library is the reference with the b2SolverTask fan-out replaced by the seam call."""
	path = b2.PKG_DIR / "libbox2d_b200.so"
	if not path.is_file():
		pytest.skip("host library not built")
	symbols = subprocess.check_output(["nm", "-D", "--defined-only", str(path)], text=True)
	assert "b2GpuSeam_SolveConstraints" in symbols
	assert "b2World_Step" in symbols
	full = subprocess.check_output(["nm", str(path)], text=True)
	assert "b2SolverTask" not in full, "the CPU solver driver must not be linked into the GPU host library"


def _allocator_round_trip():
	lib = b2.solver_lib()
	blocks = []
	# (9 MB: a block of its own, registered and released on its own -- on a GPU box on huge pages, here plain memory)
	for size in (1, 63, 64, 200, 4096, 100000, 5 << 20, 9 << 20):
		p = lib.b2GpuHostAlloc(size, 32)
		assert p and p % 64 == 0
		ctypes.memset(p, 0xAB, size)
		blocks.append((p, size))
	for p, size in blocks:
		lib.b2GpuHostFree(p, size)
	again = lib.b2GpuHostAlloc(200, 32)
	assert again in [p for p, _ in blocks]  # recycled from the free list
	lib.b2GpuHostFree(again, 200)
	big = lib.b2GpuHostAlloc(9 << 20, 32)  # (released above: a new registration)
	assert big and big % 64 == 0
	ctypes.memset(big, 0xCD, 9 << 20)
	lib.b2GpuHostFree(big, 9 << 20)


def test_pinned_allocator_round_trip():
	_allocator_round_trip()


@pytest.mark.gpu
def test_pinned_allocator_round_trip_on_the_device_box():
	"""The same through the page-locked path (b2gPinnedAlloc: registered huge pages, or cudaHostAlloc)."""
	_allocator_round_trip()


def test_island_sizes_from_labels(capture_files):
	"""b2GpuCountIslandSizes (host utility, no device): every awake body, touching contact and joint is counted once,
	in the island of its bodies."""
	import numpy as np

	for path in capture_files:
		cap = b2.Capture(path)
		d, _, bufs = cap.make_call(sizes=True)
		if "sizes" not in bufs:
			continue
		sizes = np.frombuffer(bufs["sizes"], dtype=np.int32).reshape(-1, 4)
		labels = cap.island_labels
		valid = (labels >= 0) & (labels < d.islandCount)
		assert np.array_equal(sizes[:, 0], np.bincount(labels[valid], minlength=d.islandCount))
		assert sizes[:, 1].sum() <= cap.contact_count and sizes[:, 2].sum() <= cap.joint_count
		if valid.all():
			assert sizes[:, 1].sum() == cap.contact_count
		assert (sizes[:, 3] == 0).all()


def test_no_device_means_a_loud_abort_not_a_cpu_fallback():
	"""north_star: no CPU fallback.  Without a CUDA device the first b2World_Step of the product library aborts with a message
	(box2d_b200/host/b2_gpu_seam.c, b2SeamFatal) instead of solving on the host."""
	import os
	import subprocess
	import sys

	if not (b2.PKG_DIR / "libbox2d_b200.so").is_file():
		pytest.skip("host library not built")
	code = (
		"import box2d_b200 as b2\n"
		"lib = b2.host_lib()\n"
		"w = b2.World(lib, 'small_pyramid', 1)\n"
		"w.step(3)\n"
		"print('stepped without a device')\n"
	)
	env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONPATH=str(b2.ROOT))
	proc = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
	assert proc.returncode != 0, proc.stdout
	assert "stepped without a device" not in proc.stdout
	assert "cannot create the device solver" in proc.stderr and "no CPU fallback" in proc.stderr, proc.stderr
