"""Resident mode (include/b2_gpu_solver.h, b2GpuSolverGetResidentStats): the device keeps contacts, impulses and bodies
across steps and the pack pass uploads only what differs from that.  These tests drive SEQUENCES of steps through one
solver -- each step's outputs become the next step's inputs, the way b2World_Step chains them -- with the host touching
things in between the way the reference does (recycled manifolds: only the separations move, src/physics_world.c:545-550;
re-evaluated manifolds; swap-removes in the colour arrays, src/constraint_graph.c:198-211; bodies whose velocity the user
set).  Every step is compared bit for bit with the CPU oracle on the very same inputs, and the statistics must show that
the light path was really taken.  (A contact keeps its resident record as long as it keeps its place in its colour's
array: the two contacts a swap-remove moves travel in full once.)"""
import ctypes

import numpy as np
import pytest

import box2d_b200 as b2

pytestmark = pytest.mark.gpu

MAN = 68  # b2ContactSim::manifold
POINT = (MAN + 12, MAN + 12 + 44)


@pytest.fixture(scope="module")
def oracle():
	lib = ctypes.CDLL(str(b2.ROOT / "oracle" / "liboracle.so"))
	lib.b2OracleSolverStep.restype = ctypes.c_int
	lib.b2OracleSolverStep.argtypes = [ctypes.POINTER(b2.StepDesc), ctypes.POINTER(b2.StepResult)]
	return lib


def _finalize(cap: b2.Capture, bufs) -> None:
	"""What the host does to the solver's outputs before the next step (b2FinalizeBodiesTask, src/solver.c:611-612, :632):
	the outputs become the capture's inputs, deltas reset, transient flags cleared."""
	n = cap.body_count
	st = bufs["states"].copy()
	if n:
		f = st.view(np.float32).reshape(n, 8)
		f[:, 4:8] = np.array([0.0, 0.0, 1.0, 0.0], dtype=np.float32)
		st.view(np.uint32).reshape(n, 8)[:, 3] &= np.uint32(~0x68 & 0xFFFFFFFF)
	cap.states_in = st
	cap.contacts_in = [a.copy() for a in bufs["contacts"]]
	cap.joints_in = [a.copy() for a in bufs["joints"]]


def _step_both(oracle, solver, cap, tag, **kw):
	d0, r0, want = cap.make_call(**kw)
	assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
	d, r, got = cap.make_call(**kw)
	solver.step(d, r)
	assert np.array_equal(got["states"], want["states"]), tag + ": states"
	for a, b in zip(got["contacts"], want["contacts"]):
		assert np.array_equal(a, b), tag + ": contact sims"
	for a, b in zip(got["joints"], want["joints"]):
		assert np.array_equal(a, b), tag + ": joint sims"
	assert np.array_equal(got["hit"], want["hit"]), tag + ": hit bits"
	assert np.array_equal(got["joint"], want["joint"]), tag + ": joint bits"
	assert bool(r.hasHitEvents) == bool(r0.hasHitEvents), tag
	return got


def _contacts(arr):
	return arr.reshape(-1, b2.CONTACT_SIZE)


def _mutate(cap: b2.Capture, rng, kind: str) -> int:
	"""Touch the inputs the way the host does between two steps; returns how many contacts MUST travel as full records."""
	touched = 0
	if kind == "recycle":
		# recycled manifolds: new separations, nothing else (src/physics_world.c:545-550)
		for arr in cap.contacts_in:
			if arr.size:
				c = _contacts(arr)
				for p in POINT:
					sep = c[:, p + 16:p + 20].view(np.float32)
					sep += rng.normal(0.0, 0.002, size=sep.shape).astype(np.float32)
	elif kind == "manifold":
		# re-evaluated manifolds on every third contact: anchors and normal move, impulses are matched and kept
		for arr in cap.contacts_in:
			if arr.size:
				c = _contacts(arr)[::3]
				for p in POINT:
					anchors = c[:, p:p + 16].view(np.float32)
					anchors += rng.normal(0.0, 0.01, size=anchors.shape).astype(np.float32)
				touched += c.shape[0]
	elif kind == "impulses":
		# a manifold whose point ids changed: the cached impulses are dropped (src/contact.c, b2UpdateContact)
		for arr in cap.contacts_in:
			if arr.size:
				c = _contacts(arr)[1::4]
				for p in POINT:
					c[:, p + 24:p + 32] = 0
				touched += c.shape[0]
	elif kind == "material":
		for arr in cap.contacts_in:
			if arr.size:
				c = _contacts(arr)[::5]
				c[:, 172:176].view(np.float32)[:] = rng.uniform(0.1, 0.9, size=(c.shape[0], 1)).astype(np.float32)
				touched += c.shape[0]
	elif kind == "swap":
		# swap-remove churn: the first and the last contact of every colour trade places (same ids, new slots)
		for arr in cap.contacts_in:
			c = _contacts(arr) if arr.size else None
			if c is not None and c.shape[0] >= 2:
				first = c[0].copy()
				c[0] = c[-1]
				c[-1] = first
				touched += 2  # new at their homes
	elif kind == "velocity":
		n = cap.body_count
		if n:
			st = cap.states_in.view(np.float32).reshape(n, 8)
			st[::4, 0:3] += rng.normal(0.0, 0.3, size=st[::4, 0:3].shape).astype(np.float32)
	elif kind == "force":
		n = cap.body_count
		if n:
			sims = cap.sims.reshape(n, b2.SIM_SIZE)
			sims[::6, 48:56].view(np.float32)[:] += rng.normal(0.0, 2.0, size=(sims[::6].shape[0], 2)).astype(np.float32)
	elif kind == "deltas":
		# not something the reference does (finalize resets them), but the C-ABI takes any state
		n = cap.body_count
		if n:
			st = cap.states_in.view(np.float32).reshape(n, 8)
			st[::7, 4:6] += np.float32(0.001)
	elif kind != "none":
		raise ValueError(kind)
	return touched


SEQUENCE = ["none", "recycle", "none", "manifold", "recycle", "swap", "velocity", "impulses", "material", "force", "deltas", "recycle", "none"]


@pytest.mark.parametrize("islands", [True, False], ids=["island-kernels", "grid-kernel"])
def test_chained_steps_match_the_oracle(oracle, capture_files, islands):
	rng = np.random.default_rng(7)
	light_steps = 0
	for path in capture_files:
		cap = b2.Capture(path)
		if cap.contact_count == 0:
			continue
		with b2.GpuSolver() as solver:
			for step, kind in enumerate(SEQUENCE):
				must_full = _mutate(cap, rng, kind)
				tag = f"{path.name} step {step} after {kind!r}"
				got = _step_both(oracle, solver, cap, tag, islands=islands)
				stats = solver.resident_stats()
				assert stats is not None, tag + ": not a resident step"
				full, dirty = stats
				if step == 0:
					assert full == cap.contact_count and dirty == cap.body_count, tag + ": a cold step sends everything"
				else:
					if kind == "impulses":
						assert full <= must_full, tag  # (impulses that were zero already do not make a contact dirty)
					else:
						assert full == must_full, tag + f": {full} contacts travelled in full, {must_full} were touched"
					light_steps += 1 if full == 0 else 0
					if kind in ("none", "recycle", "swap", "manifold", "impulses", "material"):
						assert dirty == 0, tag + f": {dirty} bodies re-uploaded although the host did not touch any"
					if kind in ("velocity", "force", "deltas"):
						assert 0 < dirty < max(2, cap.body_count), tag
				_finalize(cap, got)
			if islands and "small_pyramid_030" in path.name:
				# two of the steps follow a step that sent no contact in full and send none themselves, with every body in the
				# bin it was in: they run on the lists the step before left on the device (no partition kernel)
				assert solver.list_reuse_count() >= 2, solver.list_reuse_count()
			if not islands:
				assert solver.list_reuse_count() == 0
	assert light_steps > 0


def test_clusters_run_on_the_previous_steps_lists_too(oracle, capture_files, monkeypatch):
	"""The same chain of steps with every bin shared by a cluster of two blocks (b2gPartitionKernel + b2gClusterIslandKernel): the
	steady steps of the chain skip the partition kernel, everything still matches the oracle bit for bit."""
	monkeypatch.setenv("B2GPU_CLUSTER_FORCE", "2")
	rng = np.random.default_rng(11)
	cap = b2.Capture([f for f in capture_files if "small_pyramid_030" in f.name][0])
	with b2.GpuSolver() as solver:
		for step, kind in enumerate(SEQUENCE):
			_mutate(cap, rng, kind)
			got = _step_both(oracle, solver, cap, f"cluster of 2, step {step} after {kind!r}")
			assert solver.island_plan()[1] == 2, "the step did not run on clusters of two blocks"
			_finalize(cap, got)
		assert solver.list_reuse_count() >= 2, solver.list_reuse_count()


def test_contacts_leaving_and_returning(oracle, capture_files):
	"""A contact that sits out a step (stopped touching, came back) must not pick up stale impulses: its previous output
	record is two steps old."""
	rng = np.random.default_rng(11)
	path = [f for f in capture_files if "small_pyramid_030" in f.name][0]
	cap = b2.Capture(path)
	with b2.GpuSolver() as solver:
		got = _step_both(oracle, solver, cap, "warm-up")
		_finalize(cap, got)
		keep = [a.copy() for a in cap.contacts_in]
		counts = list(cap.color_counts)
		# drop the last contact of every colour for one step ...
		dropped = 0
		for i, arr in enumerate(cap.contacts_in):
			cc, jc = cap.color_counts[i]
			if cc >= 2:
				cap.contacts_in[i] = arr[: (cc - 1) * b2.CONTACT_SIZE].copy()
				cap.color_counts[i] = (cc - 1, jc)
				cd = cap.desc.colors[i] if i < cap.desc.activeColorCount else cap.desc.overflow
				cd.contactCount = cc - 1
				dropped += 1
		assert dropped > 0
		got = _step_both(oracle, solver, cap, "without them")
		assert solver.resident_stats()[0] == 0
		_finalize(cap, got)
		# ... and bring them back exactly as they were two steps ago
		for i, arr in enumerate(keep):
			cc, jc = counts[i]
			if cc >= 2:
				cap.contacts_in[i] = np.concatenate([cap.contacts_in[i], arr[(cc - 1) * b2.CONTACT_SIZE:]])
				cap.color_counts[i] = (cc, jc)
				cd = cap.desc.colors[i] if i < cap.desc.activeColorCount else cap.desc.overflow
				cd.contactCount = cc
		_mutate(cap, rng, "recycle")
		got = _step_both(oracle, solver, cap, "back again")
		assert solver.resident_stats()[0] == dropped, "exactly the returning contacts travel in full"


def test_resident_and_plain_steps_interleave(oracle, capture_files):
	"""A batch step (plain wire) in between invalidates the resident copies: the next single-world step sends everything."""
	cap = b2.Capture([f for f in capture_files if "falling_hinges_120" in f.name][0])
	other = b2.Capture([f for f in capture_files if "small_pyramid_030" in f.name][0])
	with b2.GpuSolver() as solver:
		got = _step_both(oracle, solver, cap, "first")
		_finalize(cap, got)
		got = _step_both(oracle, solver, cap, "second")
		assert solver.resident_stats()[0] == 0
		_finalize(cap, got)
		descs, results, keep = b2.make_batch([other, other])
		solver.step_batch(descs, results)
		assert solver.resident_stats() is None
		got = _step_both(oracle, solver, cap, "after the batch")
		assert solver.resident_stats() == (cap.contact_count, cap.body_count)
		_finalize(cap, got)
		# the split-phase entry points keep the copies in step too
		d, r, bufs = cap.make_call()
		d0, r0, want = cap.make_call()
		assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
		solver.upload(d)
		solver.run(r)
		solver.run(r)
		solver.download(d, r)
		assert solver.resident_stats()[0] == 0
		assert np.array_equal(bufs["states"], want["states"])
		for a, b in zip(bufs["contacts"], want["contacts"]):
			assert np.array_equal(a, b)


def test_resident_off_matches(oracle, capture_files, monkeypatch):
	monkeypatch.setenv("B2GPU_RESIDENT", "0")
	cap = b2.Capture([f for f in capture_files if "contact_zoo_130" in f.name][0])
	with b2.GpuSolver() as solver:
		for step in range(3):
			got = _step_both(oracle, solver, cap, f"plain step {step}")
			assert solver.resident_stats() is None
			_finalize(cap, got)


def _hints(cap: b2.Capture, stamp: int):
	"""b2GpuStepDesc::recycled for every contact of the capture, as the seam builds it from the narrow phase's recycle branch."""
	total = cap.contact_count
	hints = (b2.RecycledContact * max(1, total))()
	starts, counts = [], []
	k = 0
	for arr in cap.contacts_in:
		c = _contacts(arr) if arr.size else np.zeros((0, b2.CONTACT_SIZE), np.uint8)
		starts.append(k)
		counts.append(c.shape[0])
		for row in c:
			hints[k].stamp = stamp
			hints[k].contactId = int(row[0:4].view(np.int32)[0])
			hints[k].separation[0] = float(row[POINT[0] + 16:POINT[0] + 20].view(np.float32)[0])
			hints[k].separation[1] = float(row[POINT[1] + 16:POINT[1] + 20].view(np.float32)[0])
			hints[k].indexA = int(row[36:40].view(np.int32)[0])
			hints[k].indexB = int(row[40:44].view(np.int32)[0])
			k += 1
	return hints, starts, counts


def _step_with_hints(oracle, solver, cap, tag, stamp, desc_stamp=None, spoil=(), in_place=False):
	d0, r0, want = cap.make_call()
	assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
	d, r, got = cap.make_call()
	hints, starts, counts = _hints(cap, stamp)
	for k in spoil:
		hints[k].contactId += 100000  # an entry written for another contact: must be ignored
	d.recycled = ctypes.addressof(hints)
	d.recycledStamp = stamp if desc_stamp is None else desc_stamp
	for c, (a, n) in enumerate(zip(starts, counts)):
		d.recycledStart[c] = a
		d.recycledCount[c] = n
		d.recycledInPlace[c] = 1 if in_place else 0
	solver.step(d, r)
	assert np.array_equal(got["states"], want["states"]), tag + ": states"
	for a, b in zip(got["contacts"], want["contacts"]):
		assert np.array_equal(a, b), tag + ": contact sims"
	assert np.array_equal(got["hit"], want["hit"]), tag + ": hit bits"
	return got


@pytest.mark.parametrize("in_place", [False, True], ids=["by id", "in place"])
def test_recycled_hint_skips_the_comparison(oracle, capture_files, in_place):
	"""b2GpuStepDesc::recycled: contacts the narrow phase vouches for are taken without reading their record; entries with a
	stale stamp or another contact's id are ignored; the results are the oracle's either way.  In place (recycledInPlace):
	the caller also says that the colour arrays are the ones the entries were written against, and the pack pass does not
	look at the contacts at all -- their ids and body indices come from the entries."""
	rng = np.random.default_rng(3)
	for name in ("small_pyramid_030", "contact_zoo_130", "overflow_025", "falling_hinges_120"):
		cap = b2.Capture([f for f in capture_files if name in f.name][0])
		n = cap.contact_count
		with b2.GpuSolver() as solver:
			got = _step_with_hints(oracle, solver, cap, name + " cold", stamp=1, in_place=in_place)
			assert solver.vouched_contacts() == 0, "nothing can be vouched for before the device has seen it"
			_finalize(cap, got)
			_mutate(cap, rng, "recycle")
			got = _step_with_hints(oracle, solver, cap, name + " vouched", stamp=2, in_place=in_place)
			assert solver.vouched_contacts() == n and solver.resident_stats()[0] == 0
			_finalize(cap, got)
			_mutate(cap, rng, "recycle")
			got = _step_with_hints(oracle, solver, cap, name + " stale stamp", stamp=2, desc_stamp=3, in_place=in_place)
			assert solver.vouched_contacts() == 0 and solver.resident_stats()[0] == 0
			_finalize(cap, got)
			_mutate(cap, rng, "recycle")
			spoil = list(range(0, n, 3))
			got = _step_with_hints(oracle, solver, cap, name + " foreign ids", stamp=4, spoil=spoil, in_place=in_place)
			assert solver.vouched_contacts() == n - len(spoil) and solver.resident_stats()[0] == 0
			_finalize(cap, got)
			# bodies that moved in the awake set while the manifold was recycled (src/physics_world.c:497-504 refreshes the
			# indices): the entry is not enough, the record is examined -- and, its indices being part of it, sent
			moved = 0
			for arr in cap.contacts_in:
				if arr.size:
					c = _contacts(arr)
					ia = c[:, 36:40].view(np.int32)
					ib = c[:, 40:44].view(np.int32)
					swap = (ia[:, 0] >= 0) & (ib[:, 0] >= 0) & (np.arange(c.shape[0]) % 4 == 0)
					ia[swap, 0], ib[swap, 0] = ib[swap, 0].copy(), ia[swap, 0].copy()
					moved += int(swap.sum())
			got = _step_with_hints(oracle, solver, cap, name + " moved bodies", stamp=5, in_place=in_place)
			assert solver.vouched_contacts() == n - moved and solver.resident_stats()[0] == moved


def test_the_seam_vouches_for_recycled_manifolds(gpu_host_lib):
	"""Through b2World_Step: once many_pyramids has settled the narrow phase recycles every manifold
	(src/physics_world.c:508-560) and the pack pass takes all 58 000 contacts on its word."""
	with b2.World(gpu_host_lib, "many_pyramids", 8) as gpu:
		gpu.step(12)
		full, dirty, vouched = ctypes.c_int(-1), ctypes.c_int(-1), ctypes.c_int(-1)
		assert gpu_host_lib.b2GpuSeam_GetResidentStats(gpu.world_index(), ctypes.byref(full), ctypes.byref(dirty), ctypes.byref(vouched)) == 1
		contacts = sum(gpu.counters()["colorCounts"])
		assert (full.value, dirty.value, vouched.value) == (0, 0, contacts), (full.value, dirty.value, vouched.value, contacts)


def test_plain_revolute_joints_take_the_compact_records(oracle, capture_files):
	"""When every joint of a step is a plain revolute joint (no spring, motor, limit) the cluster kernel keeps 27
	words per joint instead of 64 (b2g_joint.cuh, LiteRevolute).  The plan goes by what the previous step's pack pass saw and
	the device checks: same bits on the first step (full records), on the following ones (compact), on clusters, and when a
	joint turns its motor on again (wrong guess: the step is rerun on the grid-barrier kernel)."""
	path = [f for f in capture_files if "falling_hinges_120" in f.name][0]
	for force_cluster in (None, "4"):
		import os
		if force_cluster:
			os.environ["B2GPU_CLUSTER_FORCE"] = force_cluster
		try:
			cap = b2.Capture(path)
			assert cap.joint_count > 0
			for arr in cap.joints_in:
				if arr.size:
					j = arr.reshape(-1, b2.JOINT_SIZE)
					assert (j[:, 12:16].view(np.int32) == 6).all(), "expected revolute joints only"
					j[:, 208:211] = 0  # enableSpring, enableMotor, enableLimit (include/b2gpu_layout.h, b2lRevolute)
			with b2.GpuSolver() as solver:
				for step in range(4):
					got = _step_both(oracle, solver, cap, f"plain revolutes, step {step}, cluster {force_cluster}")
					d, r, _ = cap.make_call()
					_finalize(cap, got)
				# one joint turns its motor on: the plan still counts on compact records, the device notices
				for arr in cap.joints_in:
					if arr.size:
						arr.reshape(-1, b2.JOINT_SIZE)[0, 209] = 1
						break
				d0, r0, want = cap.make_call()
				assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
				d, r, got = cap.make_call()
				solver.step(d, r)
				if force_cluster:  # (compact records are a cluster-kernel matter: one block keeps the full ones)
					assert r.gridBarriers > 0, "a joint that is not a plain revolute must send the step to the grid-barrier kernel"
				assert np.array_equal(got["states"], want["states"])
				for a, b in zip(got["joints"], want["joints"]):
					assert np.array_equal(a, b)
				_finalize(cap, got)
				got = _step_both(oracle, solver, cap, "after the wrong guess")  # full records again, island kernels
		finally:
			os.environ.pop("B2GPU_CLUSTER_FORCE", None)
