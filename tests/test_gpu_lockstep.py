"""Step-level parity: the same scene in the untouched reference (CPU solver) and in the GPU host library, stepped in
lockstep; b2World_GetStateHash (include/box2d/box2d.h:235) must agree EVERY step -- the idiom of the reference's
test/test_snapshot.c:258-283.  The hash covers transforms, velocities, contact impulses, joint impulses and the
graph layout, so agreement means the whole hot path is bit-exact, including its side outputs feeding events/sleep."""
import numpy as np
import pytest

import box2d_b200 as b2

pytestmark = pytest.mark.gpu

SCENES = [
	# scene, steps, compare every k steps
	("small_pyramid", 120, 1),
	("pyramid_soft", 60, 1),
	("pyramid_cold", 60, 1),
	("joint_zoo", 150, 1),
	("joint_zoo_cold", 60, 1),
	("contact_zoo", 240, 1),
	("overflow", 150, 1),
	("falling_hinges", 300, 1),
	("large_pyramid", 100, 5),
	("many_pyramids", 40, 5),
	("joint_grid", 100, 5),
	("rain", 300, 10),
	("tumbler", 200, 10),
	("spinner", 60, 5),
	("smash", 60, 5),
	("compounds", 40, 5),
	("washer", 40, 5),
]


def _diagnose(ref, gpu):
	"""north_star fallback metric when bits differ: worst relative error of transforms and velocities."""
	ta, tb = ref.transforms(), gpu.transforms()
	va, vb = ref.velocities(), gpu.velocities()
	if ta.shape != tb.shape:
		return f"awake sets differ: {ta.shape} vs {tb.shape}"
	scale = max(1.0, float(np.abs(ta).max()))
	vscale = max(1.0, float(np.abs(va).max()))
	return f"max |dT|/scale = {np.abs(ta - tb).max() / scale:.3e}, max |dV|/scale = {np.abs(va - vb).max() / vscale:.3e}"


@pytest.mark.parametrize("scene,steps,every", SCENES, ids=[s[0] for s in SCENES])
def test_lockstep_state_hash(ref_lib, gpu_host_lib, golden_hashes, scene, steps, every):
	with b2.World(ref_lib, scene, 4) as ref, b2.World(gpu_host_lib, scene, 4) as gpu:
		done = 0
		while done < steps:
			ref.step(every)
			gpu.step(every)
			done += every
			assert gpu.hash() == ref.hash(), f"{scene}: state hash diverged at step {done}: {_diagnose(ref, gpu)}"
		assert gpu.events() == ref.events()
		if scene in golden_hashes and golden_hashes[scene].get("steps") == steps:
			assert f"{gpu.hash():016x}" == golden_hashes[scene]["hash"]
		if scene == "falling_hinges":
			# reference test/test_determinism.c:22-23
			assert gpu.hinges_result() == (1, 274, 0xE86690F4)
		result = gpu_host_lib.b2GpuSeam_GetLastResult(gpu.world_index()).contents
		assert result.kernelLaunches >= 1


def test_multi_launch_mode_matches(ref_lib, gpu_host_lib):
	gpu_host_lib.b2GpuSeam_SetMode(1)
	try:
		with b2.World(ref_lib, "contact_zoo", 2) as ref, b2.World(gpu_host_lib, "contact_zoo", 2) as gpu:
			for _ in range(40):
				ref.step()
				gpu.step()
				assert gpu.hash() == ref.hash()
	finally:
		gpu_host_lib.b2GpuSeam_SetMode(0)
