"""Step-level parity: the same scene in the untouched reference (CPU solver) and in the GPU host library, stepped in
lockstep; b2World_GetStateHash (include/box2d/box2d.h:235) must agree EVERY step -- the idiom of the reference's
test/test_snapshot.c:258-283.  The hash covers transforms, velocities, contact impulses, joint impulses and the
graph layout, so agreement means the whole hot path is bit-exact, including its side outputs feeding events/sleep."""
import ctypes

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
	# the host changes the world between steps through the public API (box2d_b200/host/b2h_harness.c, b2hStepMutator): gravity,
	# velocities, masses, friction, motors, tuning, warm starting, teleports, body types, destroy / create, sub-step count
	("mutator", 150, 1),
]

# BASELINE.json's configurations at the step counts of the reference's benchmark (benchmark/main.c:149-160), compared EVERY step
FULL_SCALE = [
	("large_pyramid", 500),
	("many_pyramids", 200),
	("joint_grid", 500),
	("rain", 1000),
	("tumbler", 750),
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


@pytest.mark.parametrize("scene,steps", FULL_SCALE, ids=[s[0] for s in FULL_SCALE])
def test_lockstep_at_benchmark_scale(ref_lib, gpu_host_lib, scene, steps):
	"""The five benchmark scenes for as many steps as benchmark/main.c runs them, the state hash compared after every single
	step (rain: spawn / destroy / sleep churn to the end; tumbler: the overflow colour all the way)."""
	workers = 8
	with b2.World(ref_lib, scene, workers) as ref, b2.World(gpu_host_lib, scene, workers) as gpu:
		for step in range(1, steps + 1):
			ref.step()
			gpu.step()
			assert gpu.hash() == ref.hash(), f"{scene}: state hash diverged at step {step}: {_diagnose(ref, gpu)}"
		assert gpu.events() == ref.events()
		assert gpu.counters() == ref.counters()


def test_profile_stage_fields_come_from_the_device(gpu_host_lib):
	"""b2Profile's solver stage split (include/box2d/types.h:526-551) keeps being filled on the GPU path: device time per stage
	group, all of it inside b2Profile.constraints."""
	with b2.World(gpu_host_lib, "many_pyramids", 4) as gpu:
		gpu.step(5)
		p = gpu.profile()
		stages = [p[n] for n in b2.STAGE_NAMES if n != "applyRestitution"]
		assert all(v > 0.0 for v in stages), p
		assert sum(p[n] for n in b2.STAGE_NAMES) <= p["constraints"], p
		assert p["constraints"] <= p["solve"] <= p["step"]


@pytest.mark.parametrize("workers", [1, 3, 16])
def test_lockstep_with_other_worker_counts(ref_lib, gpu_host_lib, workers):
	"""The seam's team (box2d_b200/host/b2_gpu_seam.c): no helpers at all, an odd number, more workers than there is work for."""
	for scene, steps in (("mutator", 135), ("falling_hinges", 60), ("rain", 80)):
		with b2.World(ref_lib, scene, workers) as ref, b2.World(gpu_host_lib, scene, workers) as gpu:
			for step in range(steps):
				ref.step()
				gpu.step()
				assert gpu.hash() == ref.hash(), f"{scene} with {workers} workers: diverged at step {step + 1}"


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


def test_world_ids_are_reused_with_a_fresh_solver(ref_lib, gpu_host_lib):
	"""Device solvers follow the world's generation (box2d_b200/host/b2_gpu_seam.c): a world created in a destroyed world's
	slot starts from a fresh solver -- no buffers, resident copies or planner state of its predecessor."""
	gpu_host_lib.b2GpuSeam_HasSolver.restype = ctypes.c_int
	gpu_host_lib.b2GpuSeam_HasSolver.argtypes = [ctypes.c_int]
	seen = []
	for scene in ("small_pyramid", "joint_zoo", "small_pyramid"):
		with b2.World(ref_lib, scene, 2) as ref, b2.World(gpu_host_lib, scene, 2) as gpu:
			for _ in range(25):
				ref.step()
				gpu.step()
				assert gpu.hash() == ref.hash(), scene
			seen.append((gpu.world_index(), gpu_host_lib.b2GpuSeam_HasSolver(gpu.world_index())))
	assert seen[0][0] == seen[1][0] == seen[2][0], "the world id was not reused"
	assert len({tag for _, tag in seen}) == 3 and all(tag > 0 for _, tag in seen), seen
