"""Pin the oracle: oracle/_ref (the untouched reference, gcc-built from /root/reference) must reproduce the golden
vectors of the reference's own tests and the measured state hashes of SURVEY.md section 8c."""
import pytest

import box2d_b200 as b2


def test_falling_hinges_determinism_golden(ref_lib):
	"""reference test/test_determinism.c:22-23: sleepStep 274, transform hash 0xE86690F4, for any worker count."""
	for workers in (1, 3, 8):
		with b2.World(ref_lib, "falling_hinges", workers) as w:
			w.step(300)
			done, sleep_step, h = w.hinges_result()
		assert done == 1 and sleep_step == 274 and h == 0xE86690F4


def test_hello_world_like_settling(ref_lib):
	"""reference test/test_world.c:94-96 idiom: a stack settles and stays finite."""
	with b2.World(ref_lib, "small_pyramid", 1) as w:
		w.step(90)
		t = w.transforms()
	assert t.shape[0] == 55
	assert abs(t[:, 1].min() - 0.5) < 0.02  # bottom row rests on the ground (half extent 0.5)


@pytest.mark.parametrize("scene", ["small_pyramid", "joint_zoo", "contact_zoo", "overflow", "pyramid_soft", "pyramid_cold",
								   "joint_zoo_cold", "large_pyramid", "many_pyramids", "joint_grid", "tumbler"])
def test_state_hash_goldens(ref_lib, golden_hashes, scene):
	g = golden_hashes[scene]
	with b2.World(ref_lib, scene, 2) as w:
		w.step(g["steps"])
		assert f"{w.hash():016x}" == g["hash"]


def test_worker_count_invariance(ref_lib):
	"""reference test/test_snapshot.c:285-314: 1 vs 4 workers give identical hashes every step."""
	with b2.World(ref_lib, "contact_zoo", 1) as a, b2.World(ref_lib, "contact_zoo", 4) as b:
		for _ in range(60):
			a.step()
			b.step()
			assert a.hash() == b.hash()
