"""Pin the CPU oracle (oracle/liboracle.so, the plain-C restatement of the solver) against the reference itself:
for every golden capture -- the exact inputs of one solver step of the untouched reference and the outputs its CPU
solver produced -- the oracle must reproduce the outputs bit for bit.  This is what makes `oracle/` a trustworthy
checker for inputs that have no captured answer (random perturbations, full-size scenes)."""
import ctypes

import numpy as np
import pytest

import box2d_b200 as b2

ROOT = b2.ROOT


@pytest.fixture(scope="module")
def oracle():
	path = ROOT / "oracle" / "liboracle.so"
	if not path.is_file():
		from tools import buildlib
		buildlib.build_oracle_lib()
	lib = ctypes.CDLL(str(path))
	lib.b2OracleSolverStep.restype = ctypes.c_int
	lib.b2OracleSolverStep.argtypes = [ctypes.POINTER(b2.StepDesc), ctypes.POINTER(b2.StepResult)]
	return lib


def test_oracle_reproduces_reference_captures(oracle, capture_files):
	assert len(capture_files) >= 10
	for path in capture_files:
		cap = b2.Capture(path)
		desc, result, bufs = cap.make_call()
		assert oracle.b2OracleSolverStep(ctypes.byref(desc), ctypes.byref(result)) == 0
		assert np.array_equal(bufs["states"], cap.states_out), f"{path.name}: body states"
		for i, (got, want) in enumerate(zip(bufs["contacts"], cap.contacts_out)):
			assert np.array_equal(got, want), f"{path.name}: contact sims of colour slot {i}"
		for i, (got, want) in enumerate(zip(bufs["joints"], cap.joints_out)):
			assert np.array_equal(got, want), f"{path.name}: joint sims of colour slot {i}"
		assert np.array_equal(bufs["hit"][: cap.hit_bits.size], cap.hit_bits), f"{path.name}: hit bits"
		assert np.array_equal(bufs["joint"][: cap.joint_bits.size], cap.joint_bits), f"{path.name}: joint bits"
		assert bool(result.hasHitEvents) == bool(cap.has_hit_events)


def test_captures_cover_the_path(capture_files):
	"""The fixtures must reach every part of the hot path: all joint types, overflow contacts and joints, restitution,
	rolling resistance, kinematic bodies, hit events, joint events, 1- and 2-point manifolds."""
	joint_types, overflow_contacts, overflow_joints = set(), 0, 0
	restitution = rolling = hits = joint_events = one_point = two_point = 0
	for path in capture_files:
		cap = b2.Capture(path)
		for arr in cap.joints_in:
			if arr.size:
				types = arr.reshape(-1, b2.JOINT_SIZE)[:, 12:16].copy().view(np.int32).ravel()
				joint_types.update(int(t) for t in types)
		overflow_contacts += cap.color_counts[-1][0]
		overflow_joints += cap.color_counts[-1][1]
		for arr in cap.contacts_in:
			if arr.size:
				c = arr.reshape(-1, b2.CONTACT_SIZE)
				restitution += int((c[:, 176:180].copy().view(np.float32) != 0).sum())
				rolling += int((c[:, 180:184].copy().view(np.float32) != 0).sum())
				pc = c[:, 168:172].copy().view(np.int32).ravel()
				one_point += int((pc == 1).sum())
				two_point += int((pc == 2).sum())
		hits += int(np.unpackbits(cap.hit_bits.view(np.uint8)).sum())
		joint_events += int(np.unpackbits(cap.joint_bits.view(np.uint8)).sum())
	assert joint_types == set(range(9)), joint_types  # distance filter motor mover pogo prismatic revolute weld wheel
	assert overflow_contacts > 0 and overflow_joints > 0
	assert restitution > 0 and rolling > 0 and one_point > 0 and two_point > 0
	assert hits > 0 and joint_events > 0


def test_oracle_is_not_linked_into_the_product():
	"""The product must never route through the oracle (or any CPU solver)."""
	import subprocess
	for name in ("libb2gpusolver.so", "libbox2d_b200.so"):
		path = b2.PKG_DIR / name
		if path.is_file():
			syms = subprocess.check_output(["nm", "-D", str(path)], text=True)
			assert "b2OracleSolverStep" not in syms
			needed = subprocess.check_output(["readelf", "-d", str(path)], text=True)
			assert "liboracle" not in needed and "libbox2d_ref" not in needed
