"""Deferred contact impulses (include/b2_gpu_solver.h, b2GpuSolverSetDeferredImpulses; box2d_b200/host/b2_gpu_seam.c).

The reference writes every contact's impulses into its manifold at the end of every step (b2StoreImpulsesTask,
src/contact_solver.c:2293-2320); the GPU host library leaves them in the solver's page-locked output arena and writes a
manifold only when something is going to read it: the narrow phase before it re-evaluates the manifold, the pack pass,
island sleep, the contact-data / snapshot / state-hash API, hit events.  To the application nothing may look different:
these tests read the manifolds at awkward moments -- or not at all for a long time -- and compare with the untouched
reference stepping the same scene."""
import ctypes
import os

import numpy as np
import pytest

import box2d_b200 as b2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
	lib = ctypes.CDLL(str(b2.ROOT / "oracle" / "liboracle.so"))
	lib.b2OracleSolverStep.restype = ctypes.c_int
	lib.b2OracleSolverStep.argtypes = [ctypes.POINTER(b2.StepDesc), ctypes.POINTER(b2.StepResult)]
	return lib


def test_impulses_stay_in_the_arena_until_somebody_reads_a_manifold(ref_lib, gpu_host_lib):
	"""A settled pile: every manifold is recycled, nobody reads one -- nothing is written into the host's manifolds for 80
	steps.  Then the application asks for contact data (b2Shape_GetContactData on every shape): what it sees is what the
	reference's application sees, and that one question was the only flush."""
	with b2.World(ref_lib, "small_pyramid", 4) as ref, b2.World(gpu_host_lib, "small_pyramid", 4) as gpu:
		ref.step(80)
		gpu.step(80)
		pending, flushes = gpu.deferred_stats()
		assert pending and flushes == 0, (pending, flushes)
		want, count = ref.contact_checksum()
		got, got_count = gpu.contact_checksum()
		assert count > 100 and got_count == count
		assert got == want, "b2Shape_GetContactData reports other impulses than the reference's"
		pending, flushes = gpu.deferred_stats()
		assert not pending and flushes == 1, (pending, flushes)
		# ... and the world goes on as if nothing had happened
		ref.step(20)
		gpu.step(20)
		assert gpu.deferred_stats() == (True, 1)
		assert gpu.hash() == ref.hash()
		assert gpu.deferred_stats() == (False, 2)


@pytest.mark.parametrize("scene,steps", [("falling_hinges", 300), ("rain", 240), ("contact_zoo", 200), ("mutator", 150), ("tumbler", 120)])
def test_nobody_looks_until_the_end(ref_lib, gpu_host_lib, scene, steps):
	"""The whole run without a single reader from outside: islands fall asleep and wake up (falling_hinges: the reference's
	determinism golden, test/test_determinism.c:22-23), bodies are destroyed and created (rain, mutator), manifolds are
	re-evaluated, recycled, re-coloured, hit events fire (contact_zoo).  One comparison at the end."""
	with b2.World(ref_lib, scene, 4) as ref, b2.World(gpu_host_lib, scene, 4) as gpu:
		ref.step(steps)
		gpu.step(steps)
		assert gpu.events() == ref.events()
		assert gpu.contact_checksum() == ref.contact_checksum()
		assert gpu.hash() == ref.hash()
		if scene == "falling_hinges":
			assert gpu.hinges_result() == (1, 274, 0xE86690F4)


def test_snapshot_and_restore_with_impulses_pending(ref_lib, gpu_host_lib):
	"""b2World_Snapshot (include/box2d/box2d.h:316) serializes the manifolds: it must see the pending impulses.
	b2World_Restore (:329) replaces the host's contacts under the device's feet: what the device kept from the step before must
	not leak into the restored world.  Both worlds replay the same 40 steps from the same snapshot and agree every step."""
	for scene in ("large_pyramid", "joint_zoo"):
		with b2.World(ref_lib, scene, 4) as ref, b2.World(gpu_host_lib, scene, 4) as gpu:
			ref.step(30)
			gpu.step(30)
			pending, flushes = gpu.deferred_stats()
			snap_ref, snap_gpu = ref.snapshot(), gpu.snapshot()
			assert gpu.deferred_stats() == (False, flushes + (1 if pending else 0))
			assert snap_gpu[1] == snap_ref[1] == 30
			ref.step(40)
			gpu.step(40)
			first = ref.hash()
			assert gpu.hash() == first
			ref.restore(snap_ref)
			gpu.restore(snap_gpu)
			assert gpu.hash() == ref.hash(), "restored states differ"
			for step in range(40):
				ref.step()
				gpu.step()
				if step % 7 == 0:
					assert gpu.hash() == ref.hash(), f"{scene}: diverged {step + 1} steps after the restore"
			assert ref.hash() == first, "the reference itself does not replay its snapshot"
			assert gpu.hash() == first


def test_without_deferral_every_step_writes_the_manifolds(ref_lib, gpu_host_lib):
	"""B2GPU_DEFER=0 (read when a world's device solver is created): the seam stores the impulses at the end of every step
	like the reference; same results, nothing ever pending."""
	os.environ["B2GPU_DEFER"] = "0"
	try:
		with b2.World(ref_lib, "rain", 4) as ref, b2.World(gpu_host_lib, "rain", 4) as gpu:
			for _ in range(12):
				ref.step(10)
				gpu.step(10)
				assert gpu.deferred_stats() == (False, 0)
				assert gpu.hash() == ref.hash()
	finally:
		os.environ.pop("B2GPU_DEFER", None)


def _finalize_bodies(cap: b2.Capture, states) -> None:
	n = cap.body_count
	st = states.copy()
	if n:
		f = st.view(np.float32).reshape(n, 8)
		f[:, 4:8] = np.array([0.0, 0.0, 1.0, 0.0], dtype=np.float32)
		st.view(np.uint32).reshape(n, 8)[:, 3] &= np.uint32(~0x68 & 0xFFFFFFFF)
	cap.states_in = st


def _new_separations(cap: b2.Capture, rng) -> None:
	"""recycled manifolds: new separations, nothing else (src/physics_world.c:545-550)"""
	man = 68
	for arr in cap.contacts_in:
		if arr.size:
			c = arr.reshape(-1, b2.CONTACT_SIZE)
			for p in (man + 12, man + 12 + 44):
				sep = c[:, p + 16:p + 20].view(np.float32)
				sep += rng.normal(0.0, 0.002, size=sep.shape).astype(np.float32)


@pytest.mark.parametrize("name", ["small_pyramid_030", "overflow_025", "falling_hinges_120", "contact_zoo_040"])
def test_c_abi_chain_of_steps_with_stale_manifolds(oracle, capture_files, name):
	"""The C-ABI by itself.  A chain of steps through one solver with deferred impulses: the host's contact arrays are handed
	and joint arrays are handed back to the next step exactly as the previous step left them -- WITHOUT the impulses it
	computed -- the way the seam does for contacts and joints nobody has read.  The oracle runs the same chain with every impulse stored.  Bodies and joints agree after
	every step (so the device warm-started from the right impulses); the manifolds agree once they are materialized, and
	whatever the pack pass had to read in full it materialized by itself."""
	path = [f for f in capture_files if name in f.name][0]
	cap_o, cap_g = b2.Capture(path), b2.Capture(path)
	rng_o, rng_g = np.random.default_rng(5), np.random.default_rng(5)
	with b2.GpuSolver() as solver:
		solver.set_deferred(True)
		got = None
		for step in range(6):
			d0, r0, want = cap_o.make_call()
			assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
			d, r, got = cap_g.make_call()
			solver.step(d, r)
			assert np.array_equal(got["states"], want["states"]), f"step {step}: states"
			assert np.array_equal(got["joint"], want["joint"]), f"step {step}: joint bits"
			assert bool(r.hasHitEvents) == bool(r0.hasHitEvents)
			if cap_g.contact_count > 0:
				assert solver.deferred_pending()
			if step == 3:
				# somebody reads the manifolds in the middle of the run
				solver.materialize(d, got["contacts"], r, joint_arrays=got["joints"])
				for a, b in zip(got["contacts"], want["contacts"]):
					assert np.array_equal(a, b), f"step {step}: contact sims after materialize"
				for a, b in zip(got["joints"], want["joints"]):
					assert np.array_equal(a, b), f"step {step}: joint sims after materialize"
				assert np.array_equal(got["hit"], want["hit"])
				assert not solver.deferred_pending()
			# the next step's inputs: the oracle's chain has everything, the device's chain has stale manifolds
			_finalize_bodies(cap_o, want["states"])
			cap_o.contacts_in = [a.copy() for a in want["contacts"]]
			cap_o.joints_in = [a.copy() for a in want["joints"]]
			_finalize_bodies(cap_g, got["states"])
			cap_g.contacts_in = [a.copy() for a in got["contacts"]]
			cap_g.joints_in = [a.copy() for a in got["joints"]]
			_new_separations(cap_o, rng_o)
			_new_separations(cap_g, rng_g)
		d0, r0, want = cap_o.make_call()
		d, r, got = cap_g.make_call()
		# (make_call copied the inputs: materialize into the copies the last step's descriptor would have pointed at)
		assert solver.materialize(d, got["contacts"], joint_arrays=got["joints"]) == cap_g.contact_count
		for a, b in zip(got["contacts"], want["contacts"]):
			assert np.array_equal(a, b), "contact sims at the end of the chain"
		for a, b in zip(got["joints"], want["joints"]):
			assert np.array_equal(a, b), "joint sims at the end of the chain"


@pytest.mark.parametrize("knob", ["B2GPU_DIRECT_OUT", "B2GPU_KEEP_LISTS", "B2GPU_RESIDENT", "B2GPU_PDL"])
def test_each_shortcut_can_be_turned_off(ref_lib, gpu_host_lib, knob):
	"""The shortcuts of the steady state -- body states stored straight into mapped host memory, steps that run on the previous
	step's bin lists, the resident copies, the programmatic dependent launch -- are optimisations, not semantics: with any one of
	them off (the knobs are read when a world's device solver is created) the results are the same bits."""
	os.environ[knob] = "0"
	try:
		for scene, steps, every in (("many_pyramids", 24, 6), ("contact_zoo", 60, 5), ("falling_hinges", 40, 8)):
			with b2.World(ref_lib, scene, 4) as ref, b2.World(gpu_host_lib, scene, 4) as gpu:
				for _ in range(steps // every):
					ref.step(every)
					gpu.step(every)
					assert gpu.hash() == ref.hash(), f"{scene} with {knob}=0"
	finally:
		os.environ.pop(knob, None)


@pytest.mark.parametrize("env", [{"B2GPU_HUGE_PAGES": "0"}, {"B2GPU_HUGE_PAGES": "1"}, {"B2GPU_EAGER_TIMERS": "1"}])
def test_process_wide_knobs_do_not_change_the_result(ref_lib, env):
	"""Where the host arrays live (page-locked blocks from cudaHostAlloc instead of registered huge pages, B2GPU_HUGE_PAGES) and
	who reads the kernels' CUDA-event time (B2GPU_EAGER_TIMERS) are read once per process: a fresh process per setting steps
	two scenes in lockstep with the reference, with the pinned allocator installed as bench.py does."""
	import subprocess
	import sys

	code = (
		"import ctypes\n"
		"import box2d_b200 as b2\n"
		"ref = b2._bind_harness(ctypes.CDLL('oracle/_ref/libbox2d_ref.so'))\n"
		"gpu = b2.host_lib(); gpu.b2GpuSeam_InstallPinnedAllocator()\n"
		"for scene, steps, every in (('many_pyramids', 18, 6), ('joint_zoo', 40, 8)):\n"
		"    with b2.World(ref, scene, 4) as r, b2.World(gpu, scene, 4) as g:\n"
		"        for _ in range(steps // every):\n"
		"            r.step(every); g.step(every)\n"
		"            assert g.hash() == r.hash(), scene\n"
		"        assert g.profile()['constraints'] > 0.0\n"
		"print('lockstep ok')\n"
	)
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	out = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
	assert out.returncode == 0 and "lockstep ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_two_worlds_side_by_side_keep_their_impulses_apart(ref_lib, gpu_host_lib):
	"""Every world has its own device solver, output arenas and pending impulses; stepping two worlds alternately and reading
	one of them must not disturb the other.  Destroying a world with impulses pending and creating another in its slot starts
	from a clean slate."""
	with b2.World(ref_lib, "small_pyramid", 2) as ref_a, b2.World(gpu_host_lib, "small_pyramid", 2) as gpu_a:
		with b2.World(ref_lib, "contact_zoo", 2) as ref_b, b2.World(gpu_host_lib, "contact_zoo", 2) as gpu_b:
			for round_ in range(12):
				for w in (ref_a, gpu_a, ref_b, gpu_b):
					w.step(5)
				if round_ % 3 == 0:
					assert gpu_b.hash() == ref_b.hash()
					assert gpu_a.deferred_stats()[0], "reading world B must not flush world A"
				if round_ % 4 == 3:
					assert gpu_a.contact_checksum() == ref_a.contact_checksum()
			assert gpu_a.hash() == ref_a.hash() and gpu_b.hash() == ref_b.hash()
			index_b = gpu_b.world_index()
		# world B is gone with impulses pending; its slot is taken by a new world
		gpu_a.step(3)
		ref_a.step(3)
		with b2.World(ref_lib, "joint_zoo", 2) as ref_c, b2.World(gpu_host_lib, "joint_zoo", 2) as gpu_c:
			assert gpu_c.world_index() == index_b
			for _ in range(6):
				ref_c.step(5)
				gpu_c.step(5)
				assert gpu_c.hash() == ref_c.hash()
		assert gpu_a.hash() == ref_a.hash()


def test_joint_reactions_come_out_of_the_arena_joint_by_joint(ref_lib, gpu_host_lib):
	"""The joints' accumulated impulses are deferred like the contacts': b2Joint_GetConstraintForce / Torque and the per-type
	getters fetch ONE joint's record when they are asked (b2GetJointSimCheckType interposed), not the world's -- the values
	are the reference's bits, and asking is not a flush."""
	with b2.World(ref_lib, "mutator", 4) as ref, b2.World(gpu_host_lib, "mutator", 4) as gpu:
		for steps in (30, 10, 25):  # (the scene turns motors and springs on at step 32)
			ref.step(steps)
			gpu.step(steps)
			pending, flushes = gpu.deferred_stats()
			assert pending
			want, got = ref.joint_reactions(), gpu.joint_reactions()
			assert want.size >= 12 and np.array_equal(want.view(np.uint32), got.view(np.uint32)), (want, got)
			assert gpu.deferred_stats() == (True, flushes), "a joint getter flushed the whole world"
		assert gpu.hash() == ref.hash()
