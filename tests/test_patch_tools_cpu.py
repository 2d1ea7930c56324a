"""The build-time hooks (tools/patch_solver.py, tools/patch_collide.py) and the link-time interposition (tools/buildlib.py):
each hook lands exactly once in the generated translation unit, a reference source the patches were not written against is
refused, and every interposed reference function is present twice in the product library -- the seam's wrapper under the
reference's name, the reference's own definition as b2Ref_*.  No GPU; needs the reference tree (skipped without it).
"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tools import buildlib  # noqa: E402

needs_reference = pytest.mark.skipif(not buildlib.reference_available(), reason="/root/reference absent")


def _run_tool(tool: str, *args: str) -> subprocess.CompletedProcess:
	return subprocess.run([sys.executable, str(ROOT / "tools" / tool), *args], capture_output=True, text=True)


@needs_reference
def test_solver_patch_places_the_seam_call_and_the_island_hook_once(tmp_path):
	src = buildlib.REFERENCE / "src" / "solver.c"
	out = tmp_path / "solver_gpu.c"
	done = _run_tool("patch_solver.py", "--mode", "gpu", "--src", str(src), "--out", str(out))
	assert done.returncode == 0, done.stderr
	text = out.read_text()
	# (a declaration and a call each)
	assert text.count("\tb2GpuSeam_SolveConstraints( world, stepContext );") == 1
	assert text.count("\tb2GpuSeam_BeforeIslandSplit( world, stepContext );") == 1
	assert text.count("b2GpuSeam_") == 4
	# the region the seam replaces is gone: the generated unit is shorter than the source, and what follows the seam call is
	# the reference's own join of the island-split task
	assert len(text.splitlines()) < len(src.read_text(encoding="utf-8-sig").splitlines())
	after = text[text.index("\tb2GpuSeam_SolveConstraints( world"):]
	assert "// Finish island split" in after

	hook = tmp_path / "solver_hook.c"
	done = _run_tool("patch_solver.py", "--mode", "hook", "--src", str(src), "--out", str(hook))
	assert done.returncode == 0, done.stderr
	# the capture build keeps the reference's region and brackets it
	assert len(hook.read_text().splitlines()) > len(src.read_text(encoding="utf-8-sig").splitlines())


@needs_reference
def test_collide_patch_places_its_three_hooks_once(tmp_path):
	src = buildlib.REFERENCE / "src" / "physics_world.c"
	out = tmp_path / "physics_world_gpu.c"
	done = _run_tool("patch_collide.py", "--src", str(src), "--out", str(out))
	assert done.returncode == 0, done.stderr
	text = out.read_text()
	for hook in ("\tb2GpuSeam_BeginCollide( world, context, contactCount );", "\tb2GpuSeam_ContactRecycled( world, contactIndex, contactSim );",
				 "\tb2GpuSeam_ContactReevaluated( world, workerIndex, contactIndex, contactSim );"):
		assert text.count(hook) == 1, hook
	assert text.count("b2GpuSeam_") == 6  # a declaration and a call each
	# nothing of the reference is dropped by this patch
	assert len(text.splitlines()) > len(src.read_text(encoding="utf-8-sig").splitlines())


@needs_reference
@pytest.mark.parametrize("tool,source,extra", [("patch_solver.py", "solver.c", ("--mode", "gpu")), ("patch_collide.py", "physics_world.c", ())])
def test_a_source_the_patch_was_not_written_against_is_refused(tmp_path, tool, source, extra):
	changed = tmp_path / source
	changed.write_text((buildlib.REFERENCE / "src" / source).read_text(encoding="utf-8-sig") + "\n// upstream moved on\n")
	out = tmp_path / "out.c"
	done = _run_tool(tool, *extra, "--src", str(changed), "--out", str(out))
	assert done.returncode != 0 and "sha256" in done.stderr
	assert not out.exists()


def test_every_interposed_function_is_in_the_product_twice():
	lib = buildlib.PKG_DIR / "libbox2d_b200.so"
	if not lib.is_file():
		pytest.skip("host library not built")
	defined = subprocess.check_output(["nm", "--defined-only", str(lib)], text=True)
	names = [name for group in buildlib.INTERPOSED.values() for name in group]
	assert len(names) == len(set(names)) >= 13
	for name in names:
		assert f" T {name}\n" in defined, f"the seam's wrapper of {name}"
		assert f" T b2Ref_{name[2:]}\n" in defined, f"the reference's own {name}"
