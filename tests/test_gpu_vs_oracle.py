"""CUDA path vs the CPU oracle (oracle/liboracle.so) on seeded inputs that have no captured answer: the captures'
velocities, cached impulses, separations and material parameters are perturbed with a seeded RNG, then the same
descriptor is solved by b2OracleSolverStep and by b2GpuSolverStep (island kernel, grid-barrier kernel, per-stage
launches).  Integer/flag outputs and every float must agree bit for bit."""
import ctypes

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


def _perturb(cap: b2.Capture, rng: np.random.Generator) -> None:
	"""In-place, layout-aware perturbation of a capture's inputs (offsets: include/b2gpu_layout.h)."""
	n = cap.body_count
	if n:
		st = cap.states_in.view(np.float32).reshape(n, 8)
		st[:, 0:3] += rng.normal(0.0, 0.5, size=(n, 3)).astype(np.float32)  # v, w
	for arr in cap.contacts_in:
		if not arr.size:
			continue
		c = arr.reshape(-1, b2.CONTACT_SIZE)
		m = c.shape[0]

		def f32(offset, count=1):
			return c[:, offset:offset + 4 * count].view(np.float32)

		for p in (68 + 12, 68 + 12 + 44):
			sep = f32(p + 16)
			sep += rng.normal(0.0, 0.01, size=(m, 1)).astype(np.float32)  # separation, both signs
			imp = f32(p + 24, 2)
			imp[:, 0:1] = np.abs(imp[:, 0:1] + rng.normal(0.0, 0.2, size=(m, 1)).astype(np.float32))  # normal impulse >= 0
			imp[:, 1:2] += rng.normal(0.0, 0.1, size=(m, 1)).astype(np.float32)
		f32(172)[:] = rng.uniform(0.0, 1.0, size=(m, 1)).astype(np.float32)  # friction
		f32(176)[:] = np.where(rng.random((m, 1)) < 0.3, rng.uniform(0.1, 0.9, size=(m, 1)), 0.0).astype(np.float32)  # restitution
		f32(180)[:] = np.where(rng.random((m, 1)) < 0.3, rng.uniform(0.01, 0.3, size=(m, 1)), 0.0).astype(np.float32)  # rolling
		f32(184)[:] = np.where(rng.random((m, 1)) < 0.2, rng.normal(0.0, 1.0, size=(m, 1)), 0.0).astype(np.float32)  # tangent speed
		flags = c[:, 188:192].view(np.uint32)
		flags |= np.where(rng.random((m, 1)) < 0.5, np.uint32(0x00100000), np.uint32(0)).astype(np.uint32)  # hit events


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_perturbed_captures_gpu_equals_oracle(oracle, capture_files, seed):
	rng = np.random.default_rng(seed)
	with b2.GpuSolver() as solver:
		for path in capture_files:
			cap = b2.Capture(path)
			_perturb(cap, rng)
			d0, r0, want = cap.make_call()
			assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
			assert np.isfinite(want["states"].view(np.float32).reshape(-1, 8)[:, [0, 1, 2, 4, 5, 6, 7]]).all()
			for mode, islands in ((0, True), (0, False), (1, False)):
				solver.set_mode(mode)
				d, r, got = cap.make_call(islands=islands)
				solver.step(d, r)
				tag = f"{path.name} seed {seed} mode {mode} islands {islands}"
				assert np.array_equal(got["states"], want["states"]), tag + ": states"
				for a, b in zip(got["contacts"], want["contacts"]):
					assert np.array_equal(a, b), tag + ": contact sims"
				for a, b in zip(got["joints"], want["joints"]):
					assert np.array_equal(a, b), tag + ": joint sims"
				assert np.array_equal(got["hit"], want["hit"]), tag + ": hit bits"
				assert np.array_equal(got["joint"], want["joint"]), tag + ": joint bits"
				assert bool(r.hasHitEvents) == bool(r0.hasHitEvents), tag


def _drop_joints(cap: b2.Capture) -> None:
	"""Remove every joint of a capture: what is left of overflow_025 is a dynamic tray that ~20 overflow CONTACTS share --
	a deep sequential chain with no joints in it (the tumbler's drum in small)."""
	d = cap.desc
	for i in range(d.activeColorCount):
		d.colors[i].jointCount = 0
	d.overflow.jointCount = 0
	cap.joints_in = [np.zeros(0, np.uint8) for _ in cap.joints_in]
	cap.joints_out = [np.zeros(0, np.uint8) for _ in cap.joints_out]
	cap.color_counts = [(c, 0) for c, _ in cap.color_counts]


@pytest.mark.parametrize("blocks", [1, 2, 4])
def test_deep_overflow_chain_of_contacts(oracle, blocks, monkeypatch):
	"""The chain is walked by one warp whose lanes take turns (overflowChainWarp); in a cluster its bodies are cached in
	the first block for the duration of a pass.  Same bits as the oracle's one-after-the-other loop."""
	if blocks > 1:
		monkeypatch.setenv("B2GPU_CLUSTER_FORCE", str(blocks))
	rng = np.random.default_rng(7)
	with b2.GpuSolver() as solver:
		for round_ in range(3):
			cap = b2.Capture(b2.ROOT / "tests" / "golden" / "overflow_025.b2cap.gz")
			_drop_joints(cap)
			assert cap.color_counts[-1][0] >= 16 and cap.joint_count == 0
			if round_ > 0:
				_perturb(cap, rng)
			d0, r0, want = cap.make_call()
			assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
			d, r, got = cap.make_call(islands=True)
			solver.step(d, r)
			assert r.gridBarriers == 0, "expected the island-local kernels"
			bins, per_bin = solver.island_plan()
			assert per_bin >= blocks
			tag = f"round {round_} blocks {blocks}"
			assert np.array_equal(got["states"], want["states"]), tag + ": states"
			for a, b in zip(got["contacts"], want["contacts"]):
				assert np.array_equal(a, b), tag + ": contact sims"
			assert np.array_equal(got["hit"], want["hit"]), tag + ": hit bits"
			assert bool(r.hasHitEvents) == bool(r0.hasHitEvents), tag


@pytest.mark.parametrize("resident", ["0", "1"])
def test_contact_masses_that_differ_from_the_bodies(oracle, capture_files, resident, monkeypatch):
	"""b2ContactSim carries its own copy of the bodies' inverse masses (set when the contact enters the graph).  The wire
	format leaves them out when they equal the bodies' -- the usual case -- and uploads them when any contact differs:
	both ways must give the oracle's bits, and the second must move more bytes (compared with the resident mode off: there
	the byte count of a step also depends on what the device already holds)."""
	monkeypatch.setenv("B2GPU_RESIDENT", resident)
	with b2.GpuSolver() as solver:
		for path in capture_files:
			cap = b2.Capture(path)
			if cap.contact_count == 0:
				continue
			d, r, got = cap.make_call()
			solver.step(d, r)
			plain_bytes = int(r.h2dBytes)
			for arr in cap.contacts_in:
				if arr.size:
					c = arr.reshape(-1, b2.CONTACT_SIZE)
					inv = c[:, 52:68].view(np.float32)  # invMassA, invIA, invMassB, invIB
					inv[::2] *= np.float32(0.75)
			d0, r0, want = cap.make_call()
			assert oracle.b2OracleSolverStep(ctypes.byref(d0), ctypes.byref(r0)) == 0
			for islands in (True, False):
				d, r, got = cap.make_call(islands=islands)
				solver.step(d, r)
				tag = f"{path.name} islands {islands}"
				assert np.array_equal(got["states"], want["states"]), tag + ": states"
				for a, b in zip(got["contacts"], want["contacts"]):
					assert np.array_equal(a, b), tag + ": contact sims"
				if resident == "0":
					assert int(r.h2dBytes) > plain_bytes, tag + ": the masses were not uploaded"
