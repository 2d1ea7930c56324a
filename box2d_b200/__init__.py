"""box2d_b200 -- a B200-native Soft Step constraint solver behind Box2D's own host code.

The product is native: ``libb2gpusolver.so`` (hand-written sm_100a CUDA kernels + the C-ABI declared in
``include/b2_gpu_solver.h``) and ``libbox2d_b200.so`` (the reference's C17 host with the solve region of
``b2Solve`` replaced by the seam in ``box2d_b200/host``).  This package is only the thin Python mirror of
that C-ABI used by the tests and ``bench.py``: ctypes structures with the header's exact layout, loaders that
fail loudly when a library is missing (there is NO CPU fallback), and a reader for the oracle's capture files.
"""
from __future__ import annotations

import ctypes
import gzip
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
ROOT = PKG_DIR.parent

GRAPH_COLOR_COUNT = 24
MAX_ACTIVE_COLORS = GRAPH_COLOR_COUNT - 1
STAGE_NAMES = (
	"prepareConstraints", "integrateVelocities", "warmStart", "solveImpulses",
	"integratePositions", "relaxImpulses", "applyRestitution", "storeImpulses",
)

STATE_SIZE = 32
SIM_SIZE = 96
CONTACT_SIZE = 200
JOINT_SIZE = 252


# ---- include/b2_gpu_solver.h ---------------------------------------------------------------------------------
class Softness(ctypes.Structure):
	_fields_ = [("biasRate", ctypes.c_float), ("massScale", ctypes.c_float), ("impulseScale", ctypes.c_float)]


class ColorDesc(ctypes.Structure):
	_fields_ = [
		("contactSims", ctypes.c_void_p),
		("jointSims", ctypes.c_void_p),
		("contactCount", ctypes.c_int),
		("jointCount", ctypes.c_int),
		("colorIndex", ctypes.c_int),
		("reserved", ctypes.c_int),
	]


class StepDesc(ctypes.Structure):
	_fields_ = [
		("dt", ctypes.c_float), ("inv_dt", ctypes.c_float), ("h", ctypes.c_float), ("inv_h", ctypes.c_float),
		("subStepCount", ctypes.c_int),
		("contactSoftness", Softness), ("staticSoftness", Softness),
		("restitutionThreshold", ctypes.c_float), ("maxLinearVelocity", ctypes.c_float),
		("gravity", ctypes.c_float * 2),
		("contactSpeed", ctypes.c_float), ("contactHertz", ctypes.c_float), ("contactDampingRatio", ctypes.c_float),
		("hitEventThreshold", ctypes.c_float), ("lengthUnitsPerMeter", ctypes.c_float),
		("enableWarmStarting", ctypes.c_int), ("enableContactSoftening", ctypes.c_int),
		("states", ctypes.c_void_p), ("sims", ctypes.c_void_p), ("awakeBodyCount", ctypes.c_int),
		("activeColorCount", ctypes.c_int),
		("colors", ColorDesc * MAX_ACTIVE_COLORS),
		("overflow", ColorDesc),
		("contactIdCapacity", ctypes.c_int), ("jointIdCapacity", ctypes.c_int),
		("bodyIsland", ctypes.c_void_p), ("islandCount", ctypes.c_int), ("reserved0", ctypes.c_int),
		("islandSizes", ctypes.c_void_p),
		("recycled", ctypes.c_void_p), ("recycledStamp", ctypes.c_uint32),
		("recycledStart", ctypes.c_int * (MAX_ACTIVE_COLORS + 1)), ("recycledCount", ctypes.c_int * (MAX_ACTIVE_COLORS + 1)),
		("recycledInPlace", ctypes.c_int * (MAX_ACTIVE_COLORS + 1)),
	]


class RecycledContact(ctypes.Structure):
	_fields_ = [("stamp", ctypes.c_uint32), ("contactId", ctypes.c_int), ("separation", ctypes.c_float * 2),
				("indexA", ctypes.c_int), ("indexB", ctypes.c_int)]


class IslandSize(ctypes.Structure):
	_fields_ = [("bodyCount", ctypes.c_int), ("contactCount", ctypes.c_int), ("jointCount", ctypes.c_int), ("reserved", ctypes.c_int)]


class StepResult(ctypes.Structure):
	_fields_ = [
		("hitEventBits", ctypes.c_void_p),
		("jointEventBits", ctypes.c_void_p),
		("hasHitEvents", ctypes.c_int),
		("stageMs", ctypes.c_float * 8),
		("kernelMs", ctypes.c_float),
		("totalMs", ctypes.c_float),
		("h2dBytes", ctypes.c_uint64),
		("d2hBytes", ctypes.c_uint64),
		("kernelLaunches", ctypes.c_int),
		("gridBarriers", ctypes.c_int),
		("uploadMs", ctypes.c_float),
		("waitMs", ctypes.c_float),
		("scatterMs", ctypes.c_float),
		("h2dMs", ctypes.c_float),
	]


class SeamTotals(ctypes.Structure):
	_fields_ = [
		("kernelMs", ctypes.c_double), ("abiMs", ctypes.c_double), ("h2dBytes", ctypes.c_double), ("d2hBytes", ctypes.c_double),
		("stageMs", ctypes.c_double * 8),
		("steps", ctypes.c_longlong), ("launches", ctypes.c_longlong), ("gridBarriers", ctypes.c_longlong),
		("seamMs", ctypes.c_double), ("packMs", ctypes.c_double), ("waitMs", ctypes.c_double), ("unpackMs", ctypes.c_double),
		("beforeMs", ctypes.c_double),
	]


class NativeLibraryMissing(RuntimeError):
	pass


def _load(path: Path, what: str) -> ctypes.CDLL:
	if not path.is_file():
		raise NativeLibraryMissing(
			f"{what} not found at {path}: run `python -c 'import __graft_entry__ as g; g.build()'` "
			"(the solver is native CUDA; there is no Python or CPU fallback)"
		)
	return ctypes.CDLL(str(path), mode=os.RTLD_LOCAL | os.RTLD_NOW)


_solver_lib = None
_host_lib = None


def solver_lib() -> ctypes.CDLL:
	"""libb2gpusolver.so with the C-ABI prototypes of include/b2_gpu_solver.h bound."""
	global _solver_lib
	if _solver_lib is not None:
		return _solver_lib
	lib = _load(PKG_DIR / "libb2gpusolver.so", "CUDA solver library")
	P = ctypes.POINTER
	lib.b2GpuSolverCreate.restype = ctypes.c_void_p
	lib.b2GpuSolverCreate.argtypes = [ctypes.c_int]
	lib.b2GpuSolverDestroy.restype = None
	lib.b2GpuSolverDestroy.argtypes = [ctypes.c_void_p]
	for name in ("b2GpuSolverStep", "b2GpuSolverDownload"):
		fn = getattr(lib, name)
		fn.restype = ctypes.c_int
		fn.argtypes = [ctypes.c_void_p, P(StepDesc), P(StepResult)]
	lib.b2GpuSolverUpload.restype = ctypes.c_int
	lib.b2GpuSolverUpload.argtypes = [ctypes.c_void_p, P(StepDesc)]
	lib.b2GpuSolverRun.restype = ctypes.c_int
	lib.b2GpuSolverRun.argtypes = [ctypes.c_void_p, P(StepResult)]
	lib.b2GpuSolverUploadBatch.restype = ctypes.c_int
	lib.b2GpuSolverUploadBatch.argtypes = [ctypes.c_void_p, P(StepDesc), ctypes.c_int]
	lib.b2GpuSolverRunBatch.restype = ctypes.c_int
	lib.b2GpuSolverRunBatch.argtypes = [ctypes.c_void_p, P(StepResult)]
	for name in ("b2GpuSolverDownloadBatch", "b2GpuSolverStepBatch"):
		fn = getattr(lib, name)
		fn.restype = ctypes.c_int
		fn.argtypes = [ctypes.c_void_p, P(StepDesc), ctypes.c_int, P(StepResult)]
	lib.b2GpuSolverBeginStep.restype = ctypes.c_int
	lib.b2GpuSolverBeginStep.argtypes = [ctypes.c_void_p, P(StepDesc), P(StepResult)]
	for name in ("b2GpuSolverGetPackItemCount", "b2GpuSolverGetUnpackItemCount", "b2GpuSolverSubmit", "b2GpuSolverWait"):
		fn = getattr(lib, name)
		fn.restype = ctypes.c_int
		fn.argtypes = [ctypes.c_void_p]
	for name in ("b2GpuSolverPackRange", "b2GpuSolverUnpackRange"):
		fn = getattr(lib, name)
		fn.restype = None
		fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
	lib.b2GpuSolverEndStep.restype = ctypes.c_int
	lib.b2GpuSolverEndStep.argtypes = [ctypes.c_void_p, P(StepResult)]
	lib.b2GpuSolverSetMode.restype = ctypes.c_int
	lib.b2GpuSolverSetMode.argtypes = [ctypes.c_void_p, ctypes.c_int]
	lib.b2GpuSolverGetListReuseCount.restype = ctypes.c_int
	lib.b2GpuSolverGetListReuseCount.argtypes = [ctypes.c_void_p]
	lib.b2GpuSolverSetDeferredImpulses.restype = ctypes.c_int
	lib.b2GpuSolverSetDeferredImpulses.argtypes = [ctypes.c_void_p, ctypes.c_int]
	lib.b2GpuSolverDeferredPending.restype = ctypes.c_int
	lib.b2GpuSolverDeferredPending.argtypes = [ctypes.c_void_p]
	lib.b2GpuSolverDeferredSync.restype = ctypes.c_int
	lib.b2GpuSolverDeferredSync.argtypes = [ctypes.c_void_p]
	lib.b2GpuSolverMaterializeContacts.restype = ctypes.c_int
	lib.b2GpuSolverMaterializeContacts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, P(StepResult)]
	lib.b2GpuSolverMaterializeJoints.restype = ctypes.c_int
	lib.b2GpuSolverMaterializeJoints.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
	lib.b2GpuSolverDeferredForgetJoint.restype = None
	lib.b2GpuSolverDeferredForgetJoint.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
	lib.b2GpuSolverDeferredForget.restype = None
	lib.b2GpuSolverDeferredForget.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
	lib.b2GpuSolverDeferredDone.restype = None
	lib.b2GpuSolverDeferredDone.argtypes = [ctypes.c_void_p]
	lib.b2GpuHostAlloc.restype = ctypes.c_void_p
	lib.b2GpuHostAlloc.argtypes = [ctypes.c_size_t, ctypes.c_int]
	lib.b2GpuHostFree.restype = None
	lib.b2GpuHostFree.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
	lib.b2GpuGetLastError.restype = ctypes.c_char_p
	lib.b2GpuGetDeviceCount.restype = ctypes.c_int
	lib.b2GpuGetVersion.restype = ctypes.c_int
	lib.b2GpuSolverGetLaunchCount.restype = ctypes.c_uint64
	lib.b2GpuSolverGetLaunchCount.argtypes = [ctypes.c_void_p]
	lib.b2GpuSolverPackWork.restype = ctypes.c_int
	lib.b2GpuSolverPackWork.argtypes = [ctypes.c_void_p, ctypes.c_int]
	lib.b2GpuSolverUnpackWork.restype = ctypes.c_int
	lib.b2GpuSolverUnpackWork.argtypes = [ctypes.c_void_p, ctypes.c_int]
	lib.b2GpuCountIslandSizes.restype = ctypes.c_int
	lib.b2GpuCountIslandSizes.argtypes = [P(StepDesc), P(IslandSize)]
	lib.b2GpuSolverGetResidentStats.restype = ctypes.c_int
	lib.b2GpuSolverGetResidentStats.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 4
	lib.b2GpuSolverGetIslandPlan.restype = ctypes.c_int
	lib.b2GpuSolverGetIslandPlan.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
	_solver_lib = lib
	return lib


def _bind_harness(lib: ctypes.CDLL) -> ctypes.CDLL:
	"""Prototypes of box2d_b200/host/b2h_harness.c (linked into every host library variant)."""
	lib.b2h_create.restype = ctypes.c_int
	lib.b2h_create.argtypes = [ctypes.c_char_p, ctypes.c_int]
	lib.b2h_create_variant.restype = ctypes.c_int
	lib.b2h_create_variant.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
	lib.b2h_step_many.restype = ctypes.c_int
	lib.b2h_step_many.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int]
	lib.b2h_destroy.argtypes = [ctypes.c_int]
	lib.b2h_step.argtypes = [ctypes.c_int, ctypes.c_int]
	lib.b2h_set_substeps.argtypes = [ctypes.c_int, ctypes.c_int]
	lib.b2h_hash.restype = ctypes.c_uint64
	lib.b2h_hash.argtypes = [ctypes.c_int]
	lib.b2h_world_index.restype = ctypes.c_int
	lib.b2h_world_index.argtypes = [ctypes.c_int]
	lib.b2h_hinges_result.restype = ctypes.c_int
	lib.b2h_hinges_result.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint32)]
	lib.b2h_profile.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
	lib.b2h_profile_float_count.restype = ctypes.c_int
	lib.b2h_counters.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
	lib.b2h_event_counts.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
	lib.b2h_move_transforms.restype = ctypes.c_int
	lib.b2h_move_transforms.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
	lib.b2h_move_velocities.restype = ctypes.c_int
	lib.b2h_move_velocities.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
	lib.b2h_bench.restype = ctypes.c_float
	lib.b2h_bench.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
							  ctypes.POINTER(ctypes.c_float)]
	lib.b2h_version.restype = ctypes.c_int
	lib.b2h_contact_checksum.restype = ctypes.c_uint64
	lib.b2h_contact_checksum.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
	lib.b2h_mut_joint_reactions.restype = ctypes.c_int
	lib.b2h_mut_joint_reactions.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
	lib.b2h_snapshot.restype = ctypes.c_int
	lib.b2h_snapshot.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
	lib.b2h_restore.restype = ctypes.c_int
	lib.b2h_restore.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
	lib.b2h_step_index.restype = ctypes.c_int
	lib.b2h_step_index.argtypes = [ctypes.c_int]
	return lib


def host_lib() -> ctypes.CDLL:
	"""libbox2d_b200.so: the reference's host code + seam; b2World_Step solves on the GPU (no CPU solver path)."""
	global _host_lib
	if _host_lib is not None:
		return _host_lib
	solver_lib()  # dependency, resolved through $ORIGIN rpath as well
	lib = _bind_harness(_load(PKG_DIR / "libbox2d_b200.so", "GPU host library"))
	lib.b2GpuSeam_SetMode.argtypes = [ctypes.c_int]
	lib.b2GpuSeam_GetLastResult.restype = ctypes.POINTER(StepResult)
	lib.b2GpuSeam_GetLastResult.argtypes = [ctypes.c_int]
	lib.b2GpuSeam_GetLastDesc.restype = ctypes.POINTER(StepDesc)
	lib.b2GpuSeam_GetLastDesc.argtypes = [ctypes.c_int]
	lib.b2GpuSeam_GetTotals.restype = None
	lib.b2GpuSeam_GetTotals.argtypes = [ctypes.c_int, ctypes.POINTER(SeamTotals), ctypes.c_int]
	lib.b2GpuSeam_CreateGroup.restype = ctypes.c_int
	lib.b2GpuSeam_CreateGroup.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
	lib.b2GpuSeam_DestroyGroup.restype = None
	lib.b2GpuSeam_DestroyGroup.argtypes = [ctypes.c_int]
	lib.b2GpuSeam_GetResidentStats.restype = ctypes.c_int
	lib.b2GpuSeam_GetResidentStats.argtypes = [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3
	lib.b2GpuSeam_GetDeferredStats.restype = ctypes.c_int
	lib.b2GpuSeam_GetDeferredStats.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong)]
	lib.b2GpuSeam_FlushImpulses.restype = None
	lib.b2GpuSeam_FlushImpulses.argtypes = [ctypes.c_int]
	lib.b2GpuSeam_InstallPinnedAllocator.restype = None
	lib.b2GpuSeam_Shutdown.restype = None
	_host_lib = lib
	return lib


def step_many(lib: ctypes.CDLL, worlds, steps: int) -> None:
	"""Step the worlds concurrently, one native thread per world (b2h_step_many)."""
	handles = (ctypes.c_int * len(worlds))(*[w.handle for w in worlds])
	if lib.b2h_step_many(handles, len(worlds), steps) != 0:
		raise RuntimeError("b2h_step_many: could not start a thread per world")


class WorldGroup:
	"""Worlds of the GPU host library that are stepped as ONE batch per step (b2GpuSeam_CreateGroup)."""

	def __init__(self, lib: ctypes.CDLL, worlds):
		self.lib = lib
		self.worlds = list(worlds)
		indices = (ctypes.c_int * len(self.worlds))(*[w.world_index() for w in self.worlds])
		self.group = lib.b2GpuSeam_CreateGroup(indices, len(self.worlds))
		if self.group < 0:
			raise RuntimeError("b2GpuSeam_CreateGroup failed")

	def step(self, steps: int = 1) -> None:
		step_many(self.lib, self.worlds, steps)

	def close(self) -> None:
		if self.group >= 0:
			self.lib.b2GpuSeam_DestroyGroup(self.group)
			self.group = -1

	def __enter__(self):
		return self

	def __exit__(self, *exc):
		self.close()


class World:
	"""A scene in one of the host libraries (reference or GPU), driven through the b2h_* harness."""

	def __init__(self, lib: ctypes.CDLL, scene: str, workers: int = 1, variant: int = 0):
		self.lib = lib
		self.handle = lib.b2h_create_variant(scene.encode(), workers, variant) if variant else lib.b2h_create(scene.encode(), workers)
		if self.handle < 0:
			raise ValueError(f"unknown scene {scene!r} or no free world slot ({self.handle})")

	def step(self, n: int = 1) -> None:
		self.lib.b2h_step(self.handle, n)

	def hash(self) -> int:
		return int(self.lib.b2h_hash(self.handle))

	def counters(self) -> dict:
		out = (ctypes.c_int * 30)()
		self.lib.b2h_counters(self.handle, out)
		keys = ("bodyCount", "shapeCount", "contactCount", "jointCount", "islandCount", "awakeBodyCount")
		d = dict(zip(keys, out[:6]))
		d["colorCounts"] = list(out[6:30])
		return d

	def events(self) -> dict:
		out = (ctypes.c_int * 5)()
		self.lib.b2h_event_counts(self.handle, out)
		return dict(zip(("move", "begin", "end", "hit", "joint"), out))

	def profile(self) -> dict:
		n = self.lib.b2h_profile_float_count()
		out = (ctypes.c_float * n)()
		self.lib.b2h_profile(self.handle, out)
		names = ("step", "pairs", "collide", "solve", "solverSetup", "constraints", "prepareConstraints",
				 "integrateVelocities", "warmStart", "solveImpulses", "integratePositions", "relaxImpulses",
				 "applyRestitution", "storeImpulses", "splitIslands", "transforms", "sensorHits", "jointEvents",
				 "hitEvents", "refit", "bullets", "sleepIslands", "sensors")
		return dict(zip(names, out))

	def transforms(self, max_bodies: int = 1 << 20) -> np.ndarray:
		buf = np.zeros((max_bodies, 4), dtype=np.float32)
		n = self.lib.b2h_move_transforms(self.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), max_bodies)
		return buf[:n].copy()

	def velocities(self, max_bodies: int = 1 << 20) -> np.ndarray:
		buf = np.zeros((max_bodies, 3), dtype=np.float32)
		n = self.lib.b2h_move_velocities(self.handle, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), max_bodies)
		return buf[:n].copy()

	def bench(self, steps: int) -> dict:
		c = ctypes.c_float()
		s = ctypes.c_float()
		stages = (ctypes.c_float * 8)()
		total = self.lib.b2h_bench(self.handle, steps, ctypes.byref(c), ctypes.byref(s), stages)
		return {"wall_ms": float(total), "constraints_ms": float(c.value), "step_ms": float(s.value),
				"stages_ms": dict(zip(STAGE_NAMES, stages))}

	def contact_checksum(self):
		"""(checksum, contact count) of what b2Shape_GetContactData reports for every shape (the impulses an application sees)."""
		n = ctypes.c_int()
		return int(self.lib.b2h_contact_checksum(self.handle, ctypes.byref(n))), n.value

	def joint_reactions(self) -> np.ndarray:
		"""The mutator scene's joints as the application sees them: constraint force and torque of each, motor torque of the first."""
		buf = (ctypes.c_float * 64)()
		n = self.lib.b2h_mut_joint_reactions(self.handle, buf, 64)
		return np.array(buf[:n], dtype=np.float32)

	def snapshot(self) -> tuple:
		"""(image, step index): b2World_Snapshot."""
		size = self.lib.b2h_snapshot(self.handle, None, 0)
		buf = (ctypes.c_uint8 * size)()
		got = self.lib.b2h_snapshot(self.handle, buf, size)
		assert got == size, (got, size)
		return bytes(buf), self.lib.b2h_step_index(self.handle)

	def restore(self, snap: tuple) -> None:
		image, step_index = snap
		if self.lib.b2h_restore(self.handle, image, len(image), step_index) != 1:
			raise RuntimeError("b2World_Restore failed")

	def deferred_stats(self) -> tuple:
		"""GPU host library only: (impulses pending?, number of whole-world flushes so far)."""
		pending, flushes = ctypes.c_int(), ctypes.c_longlong()
		if self.lib.b2GpuSeam_GetDeferredStats(self.world_index(), ctypes.byref(pending), ctypes.byref(flushes)) != 1:
			raise RuntimeError("no device solver for this world yet")
		return bool(pending.value), int(flushes.value)

	def hinges_result(self):
		sleep_step = ctypes.c_int()
		h = ctypes.c_uint32()
		done = self.lib.b2h_hinges_result(self.handle, ctypes.byref(sleep_step), ctypes.byref(h))
		return done, sleep_step.value, h.value

	def world_index(self) -> int:
		return self.lib.b2h_world_index(self.handle)

	def destroy(self) -> None:
		if self.handle >= 0:
			self.lib.b2h_destroy(self.handle)
			self.handle = -1

	def __enter__(self):
		return self

	def __exit__(self, *exc):
		self.destroy()


# ---- capture files (oracle/harness/b2h_capture.c, format B2CAP003) ----------------------------------------------
class Capture:
	"""One captured solver step: the C-ABI inputs and the reference CPU solver's outputs."""

	def __init__(self, path):
		path = Path(path)
		raw = gzip.open(path, "rb").read() if path.suffix == ".gz" else path.read_bytes()
		if raw[:8] != b"B2CAP003":
			raise ValueError(f"{path}: not a B2CAP003 capture")
		desc_bytes = int.from_bytes(raw[8:12], "little")
		# the descriptor only ever grows at its end (optional hints, NULL = absent): older captures are zero-extended
		if desc_bytes > ctypes.sizeof(StepDesc) or desc_bytes < StepDesc.islandSizes.offset:
			raise ValueError(f"{path}: descriptor size {desc_bytes}, expected {ctypes.sizeof(StepDesc)}")
		self.desc = StepDesc.from_buffer_copy(raw[12:12 + desc_bytes].ljust(ctypes.sizeof(StepDesc), b"\0"))
		self._pos = 12 + desc_bytes
		self._raw = raw
		d = self.desc
		n = d.awakeBodyCount
		colors = [d.colors[c] for c in range(d.activeColorCount)] + [d.overflow]
		self.color_counts = [(c.contactCount, c.jointCount) for c in colors]

		self.states_in = self._take(n * STATE_SIZE)
		self.sims = self._take(n * SIM_SIZE)
		self.island_labels = self._take(n * 4).view(np.int32).copy()
		self.contacts_in, self.joints_in = [], []
		for cc, jc in self.color_counts:
			self.contacts_in.append(self._take(cc * CONTACT_SIZE))
			self.joints_in.append(self._take(jc * JOINT_SIZE))
		self.states_out = self._take(n * STATE_SIZE)
		self.contacts_out, self.joints_out = [], []
		for cc, jc in self.color_counts:
			self.contacts_out.append(self._take(cc * CONTACT_SIZE))
			self.joints_out.append(self._take(jc * JOINT_SIZE))
		hit_words = int.from_bytes(self._take(4).tobytes(), "little")
		self.hit_bits = self._take(hit_words * 8).view(np.uint64).copy()
		joint_words = int.from_bytes(self._take(4).tobytes(), "little")
		self.joint_bits = self._take(joint_words * 8).view(np.uint64).copy()
		self.has_hit_events = int.from_bytes(self._take(4).tobytes(), "little")
		del self._raw

	def _take(self, nbytes: int) -> np.ndarray:
		a = np.frombuffer(self._raw, dtype=np.uint8, count=nbytes, offset=self._pos).copy()
		self._pos += nbytes
		return a

	@property
	def body_count(self) -> int:
		return self.desc.awakeBodyCount

	@property
	def contact_count(self) -> int:
		return sum(c for c, _ in self.color_counts)

	@property
	def joint_count(self) -> int:
		return sum(j for _, j in self.color_counts)

	def make_call(self, islands: bool = True, sizes: bool = False):
		"""Fresh, writable copies of the inputs wired into a StepDesc + StepResult (the arrays must outlive the call).
		islands=False drops the island hint: the step is then solved by the grid-barrier kernel.  sizes=True adds the
		optional b2GpuStepDesc::islandSizes (bins packed by their real size)."""
		d = StepDesc.from_buffer_copy(bytes(self.desc))
		bufs = {
			"states": self.states_in.copy(),
			"sims": self.sims.copy(),
			"contacts": [a.copy() for a in self.contacts_in],
			"joints": [a.copy() for a in self.joints_in],
			"hit": np.zeros(max(1, (d.contactIdCapacity + 63) // 64), dtype=np.uint64),
			"joint": np.zeros(max(1, (d.jointIdCapacity + 63) // 64), dtype=np.uint64),
		}
		d.states = bufs["states"].ctypes.data
		d.sims = bufs["sims"].ctypes.data
		bufs["islands"] = self.island_labels.copy()
		if islands and self.island_labels.size:
			d.bodyIsland = bufs["islands"].ctypes.data
		else:
			d.bodyIsland = None
			d.islandCount = 0
		for i in range(d.activeColorCount + 1):
			cd = d.colors[i] if i < d.activeColorCount else d.overflow
			cd.contactSims = bufs["contacts"][i].ctypes.data if bufs["contacts"][i].size else None
			cd.jointSims = bufs["joints"][i].ctypes.data if bufs["joints"][i].size else None
		r = StepResult()
		r.hitEventBits = bufs["hit"].ctypes.data
		r.jointEventBits = bufs["joint"].ctypes.data
		if sizes and d.bodyIsland and d.islandCount > 0:
			# what the seam reads off b2Island (src/island.h:64-73), recounted from the labels by the library's host utility
			bufs["sizes"] = (IslandSize * d.islandCount)()
			if solver_lib().b2GpuCountIslandSizes(ctypes.byref(d), bufs["sizes"]) != 0:
				raise RuntimeError("b2GpuCountIslandSizes failed")
			d.islandSizes = ctypes.addressof(bufs["sizes"])
		return d, r, bufs


# byte ranges of b2ContactSim that the solver writes (include/b2gpu_layout.h): manifold.rollingImpulse and, per
# point, normalImpulse, tangentImpulse, totalNormalImpulse, normalVelocity
def make_batch(captures, islands: bool = True, sizes: bool = False):
	"""Wire a list of captures into contiguous (StepDesc * n), (StepResult * n) arrays + the buffers that back them."""
	n = len(captures)
	descs = (StepDesc * n)()
	results = (StepResult * n)()
	keep = []
	for i, cap in enumerate(captures):
		d, r, bufs = cap.make_call(islands=islands, sizes=sizes)
		ctypes.memmove(ctypes.byref(descs[i]), ctypes.byref(d), ctypes.sizeof(StepDesc))
		ctypes.memmove(ctypes.byref(results[i]), ctypes.byref(r), ctypes.sizeof(StepResult))
		keep.append(bufs)
	return descs, results, keep


def contact_output_view(contacts: np.ndarray) -> np.ndarray:
	"""[n, 9] float32 view of the solver-written fields of a b2ContactSim byte array."""
	n = contacts.size // CONTACT_SIZE
	c = contacts.reshape(n, CONTACT_SIZE)
	man = 68
	cols = [c[:, man + 8:man + 12]]
	for j in range(2):
		p = man + 12 + 44 * j
		cols.append(c[:, p + 24:p + 40])
	return np.ascontiguousarray(np.concatenate(cols, axis=1)).view(np.float32)


class GpuSolver:
	"""RAII wrapper over b2GpuSolverCreate / Destroy."""

	def __init__(self, device: int = 0, mode: int = 0):
		self.lib = solver_lib()
		self.handle = self.lib.b2GpuSolverCreate(device)
		if not self.handle:
			raise RuntimeError("b2GpuSolverCreate failed: " + self.lib.b2GpuGetLastError().decode())
		self.lib.b2GpuSolverSetMode(self.handle, mode)

	def _check(self, rc: int, what: str) -> None:
		if rc != 0:
			raise RuntimeError(f"{what} failed: " + self.lib.b2GpuGetLastError().decode())

	def set_mode(self, mode: int) -> None:
		self._check(self.lib.b2GpuSolverSetMode(self.handle, mode), "b2GpuSolverSetMode")

	def step(self, desc: StepDesc, result: StepResult) -> None:
		self._check(self.lib.b2GpuSolverStep(self.handle, ctypes.byref(desc), ctypes.byref(result)), "b2GpuSolverStep")

	def list_reuse_count(self) -> int:
		"""Steps that ran on the previous step's bin lists (b2GpuSolverGetListReuseCount)."""
		return int(self.lib.b2GpuSolverGetListReuseCount(self.handle))

	def set_deferred(self, enabled: bool) -> None:
		self._check(self.lib.b2GpuSolverSetDeferredImpulses(self.handle, 1 if enabled else 0), "b2GpuSolverSetDeferredImpulses")

	def deferred_pending(self) -> bool:
		return bool(self.lib.b2GpuSolverDeferredPending(self.handle))

	def materialize(self, desc: StepDesc, contact_arrays, result: StepResult = None, done: bool = True, joint_arrays=None) -> int:
		"""b2GpuSolverMaterializeContacts (and, with joint_arrays, b2GpuSolverMaterializeJoints) over the colours' byte arrays
		of b2ContactSim / b2JointSim (in the descriptor's order: the active colours, then the overflow colour); done = nothing
		is pending afterwards.  Returns the number of manifolds written."""
		total = 0
		for c, arr in enumerate(joint_arrays or []):
			if arr.size:
				color = desc.colors[c] if c < desc.activeColorCount else desc.overflow
				if self.lib.b2GpuSolverMaterializeJoints(self.handle, color.colorIndex, 0, arr.ctypes.data, arr.size // JOINT_SIZE) < 0:
					raise RuntimeError("b2GpuSolverMaterializeJoints failed: " + self.lib.b2GpuGetLastError().decode())
		for c, arr in enumerate(contact_arrays):
			if arr.size:
				color = desc.colors[c] if c < desc.activeColorCount else desc.overflow
				n = self.lib.b2GpuSolverMaterializeContacts(self.handle, color.colorIndex, 0, arr.ctypes.data, arr.size // CONTACT_SIZE,
															ctypes.byref(result) if result is not None else None)
				if n < 0:
					raise RuntimeError("b2GpuSolverMaterializeContacts failed: " + self.lib.b2GpuGetLastError().decode())
				total += n
		if done:
			self.lib.b2GpuSolverDeferredDone(self.handle)
		return total

	def upload(self, desc: StepDesc) -> None:
		self._check(self.lib.b2GpuSolverUpload(self.handle, ctypes.byref(desc)), "b2GpuSolverUpload")

	def run(self, result: StepResult) -> None:
		self._check(self.lib.b2GpuSolverRun(self.handle, ctypes.byref(result)), "b2GpuSolverRun")

	def download(self, desc: StepDesc, result: StepResult) -> None:
		self._check(self.lib.b2GpuSolverDownload(self.handle, ctypes.byref(desc), ctypes.byref(result)),
					"b2GpuSolverDownload")

	def step_phased(self, desc: StepDesc, result: StepResult, workers: int = 4, pipelined: bool = True) -> None:
		"""The step through the phased entry points the seam uses, with `workers` host threads (ctypes drops the GIL):
		pipelined = BeginStep, PackWork x workers, Submit, UnpackWork x workers, EndStep;
		otherwise  = BeginStep, PackRange x workers, Submit, Wait, UnpackRange x workers, EndStep."""
		import threading

		lib, h = self.lib, self.handle
		self._check(lib.b2GpuSolverBeginStep(h, ctypes.byref(desc), ctypes.byref(result)), "b2GpuSolverBeginStep")
		failures = []

		def run_all(fn_for):
			threads = [threading.Thread(target=fn_for(i)) for i in range(1, workers)]
			for t in threads:
				t.start()
			fn_for(0)()
			for t in threads:
				t.join()

		if pipelined:
			def pack(i):
				return lambda: failures.append("pack") if lib.b2GpuSolverPackWork(h, 1 if i == 0 else 0) != 0 else None

			def unpack(i):
				return lambda: failures.append("unpack") if lib.b2GpuSolverUnpackWork(h, 1 if i == 0 else 0) != 0 else None

			run_all(pack)
			self._check(lib.b2GpuSolverSubmit(h), "b2GpuSolverSubmit")
			run_all(unpack)
		else:
			n = lib.b2GpuSolverGetPackItemCount(h)
			cuts = [n * i // workers for i in range(workers + 1)]
			run_all(lambda i: (lambda: lib.b2GpuSolverPackRange(h, cuts[i], cuts[i + 1])))
			self._check(lib.b2GpuSolverSubmit(h), "b2GpuSolverSubmit")
			self._check(lib.b2GpuSolverWait(h), "b2GpuSolverWait")
			m = lib.b2GpuSolverGetUnpackItemCount(h)
			cuts = [m * i // workers for i in range(workers + 1)]
			run_all(lambda i: (lambda: lib.b2GpuSolverUnpackRange(h, cuts[i], cuts[i + 1])))
		if failures:
			raise RuntimeError(f"phased step failed in {failures}: " + lib.b2GpuGetLastError().decode())
		self._check(lib.b2GpuSolverEndStep(h, ctypes.byref(result)), "b2GpuSolverEndStep")

	def step_batch(self, descs, results) -> None:
		"""descs / results: ctypes arrays (StepDesc * n), (StepResult * n) -- b2GpuSolverStepBatch."""
		self._check(self.lib.b2GpuSolverStepBatch(self.handle, descs, len(descs), results), "b2GpuSolverStepBatch")

	def upload_batch(self, descs) -> None:
		self._check(self.lib.b2GpuSolverUploadBatch(self.handle, descs, len(descs)), "b2GpuSolverUploadBatch")

	def run_batch(self, result: StepResult) -> None:
		self._check(self.lib.b2GpuSolverRunBatch(self.handle, ctypes.byref(result)), "b2GpuSolverRunBatch")

	def download_batch(self, descs, results) -> None:
		self._check(self.lib.b2GpuSolverDownloadBatch(self.handle, descs, len(descs), results), "b2GpuSolverDownloadBatch")

	def launch_count(self) -> int:
		return int(self.lib.b2GpuSolverGetLaunchCount(self.handle))

	def island_plan(self) -> tuple[int, int]:
		"""(bins, thread blocks per bin) of the last step; (0, 0) when it was planned for the grid-barrier kernel."""
		bins, blocks = ctypes.c_int(0), ctypes.c_int(0)
		self.lib.b2GpuSolverGetIslandPlan(self.handle, ctypes.byref(bins), ctypes.byref(blocks))
		return bins.value, blocks.value

	def resident_stats(self):
		"""(full contact records, dirty bodies) of the last step, or None when it did not run in resident mode."""
		full, dirty = ctypes.c_int(0), ctypes.c_int(0)
		if self.lib.b2GpuSolverGetResidentStats(self.handle, ctypes.byref(full), ctypes.byref(dirty), None, None) == 0:
			return None
		return full.value, dirty.value

	def vouched_contacts(self) -> int:
		"""Contacts of the last step the pack pass took on the caller's word (b2GpuStepDesc::recycled)."""
		n = ctypes.c_int(0)
		self.lib.b2GpuSolverGetResidentStats(self.handle, None, None, ctypes.byref(n), None)
		return n.value

	def full_joints(self) -> int:
		"""Joints of the last step that travelled as complete 256-byte records (resident mode)."""
		n = ctypes.c_int(0)
		self.lib.b2GpuSolverGetResidentStats(self.handle, None, None, None, ctypes.byref(n))
		return n.value

	def close(self) -> None:
		if self.handle:
			self.lib.b2GpuSolverDestroy(self.handle)
			self.handle = None

	def __enter__(self):
		return self

	def __exit__(self, *exc):
		self.close()
