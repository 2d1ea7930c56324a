"""Batch of independent worlds through b2GpuSolverStepBatch: every world of the batch must come out bit-identical to
what the reference's CPU solver produced for that world alone (the captures), whatever the mix of scenes, joint types,
overflow constraints and world sizes in the batch."""
import ctypes

import numpy as np
import pytest

import box2d_b200 as b2
from test_gpu_capture import _check

pytestmark = pytest.mark.gpu

# captures that share the default step parameters (a batch must: the stage parameters are launch constants)
NAMES = ["small_pyramid_000", "falling_hinges_120", "joint_zoo_045", "overflow_025", "small_pyramid_030", "joint_zoo_000",
		 "falling_hinges_020", "overflow_001"]


def _captures(repeat=1):
	caps = [b2.Capture(b2.ROOT / "tests" / "golden" / f"{name}.b2cap.gz") for name in NAMES]
	return caps * repeat


@pytest.mark.parametrize("islands", [True, False])
def test_batch_matches_each_world(islands):
	caps = _captures(repeat=3)
	descs, results, bufs = b2.make_batch(caps, islands=islands)
	with b2.GpuSolver() as solver:
		solver.step_batch(descs, results)
	for cap, buf, res in zip(caps, bufs, results):
		_check(cap, buf, res)
	if islands:
		assert results[0].gridBarriers == 0, "the island kernel should have solved the batch"


def test_batch_split_phase_and_many_worlds():
	caps = [b2.Capture(b2.ROOT / "tests" / "golden" / "small_pyramid_030.b2cap.gz")] * 700
	descs, results, bufs = b2.make_batch(caps)
	with b2.GpuSolver() as solver:
		solver.upload_batch(descs)
		r = b2.StepResult()
		solver.run_batch(r)
		solver.run_batch(r)
		solver.download_batch(descs, results)
		assert r.gridBarriers == 0
	for i in (0, 1, 349, 698, 699):
		_check(caps[i], bufs[i], results[i])


def test_many_worlds_with_island_sizes_run_in_waves():
	"""Islands that are a sizeable fraction of a block, more bins than SMs: the planner packs by the real sizes and still
	has to find a bin count whose fullest bin fits (b2gPlanBins retries with more bins)."""
	caps = [b2.Capture(b2.ROOT / "tests" / "golden" / "small_pyramid_030.b2cap.gz")] * 2500
	descs, results, bufs = b2.make_batch(caps, sizes=True)
	with b2.GpuSolver() as solver:
		solver.step_batch(descs, results)
		bins, blocks = solver.island_plan()
		assert results[0].gridBarriers == 0 and bins > 148 and blocks == 1, "the island kernel should have solved the batch in waves"
	for i in (0, 1, 1250, 2498, 2499):
		_check(caps[i], bufs[i], results[i])


def test_batch_rejects_mixed_step_parameters():
	caps = [b2.Capture(b2.ROOT / "tests" / "golden" / "small_pyramid_030.b2cap.gz"),
			b2.Capture(b2.ROOT / "tests" / "golden" / "pyramid_cold_003.b2cap.gz")]  # warm starting off
	descs, results, bufs = b2.make_batch(caps)
	with b2.GpuSolver() as solver:
		with pytest.raises(RuntimeError):
			solver.step_batch(descs, results)


def test_large_batch_through_the_pipelined_host_passes():
	"""700 worlds = ~140 000 pack items: the one-call entry point splits the two host passes over several threads that claim
	blocks concurrently while the pump thread moves finished prefixes over PCIe (b2GpuSolverPackWork / UnpackWork)."""
	caps = [b2.Capture(b2.ROOT / "tests" / "golden" / f"{name}.b2cap.gz") for name in ("small_pyramid_030", "falling_hinges_120")] * 350
	descs, results, bufs = b2.make_batch(caps)
	with b2.GpuSolver() as solver:
		solver.step_batch(descs, results)
	for cap, buf, res in zip(caps, bufs, results):
		_check(cap, buf, res)


@pytest.mark.parametrize("blocks", [2, 8])
def test_batch_on_clusters(blocks, monkeypatch):
	"""Several bins, each shared by a cluster: the colours are dealt out evenly over the blocks (owner lists need a
	single bin), joints and contacts use counted stores."""
	monkeypatch.setenv("B2GPU_CLUSTER_FORCE", str(blocks))
	caps = _captures(repeat=2)
	descs, results, bufs = b2.make_batch(caps)
	with b2.GpuSolver() as solver:
		solver.step_batch(descs, results)
		bins, per_bin = solver.island_plan()
		assert bins > 1 and per_bin >= blocks
	for cap, buf, res in zip(caps, bufs, results):
		_check(cap, buf, res)


def test_group_of_live_worlds_steps_as_one_batch(ref_lib, gpu_host_lib):
	"""World-level batched step (b2GpuSeam_CreateGroup): 96 live worlds, each its own slightly different pile (the offset
	is derived from the world's index), stepped concurrently -- broad phase, narrow phase and finalize per world on the host,
	ONE b2GpuSolverStepBatch per round of steps.  Every world is compared with the reference stepping the same world."""
	count, rounds, per_round = 96, 12, 10
	refs = [b2.World(ref_lib, "small_pyramid", 1, variant=1 + i) for i in range(count)]
	gpus = [b2.World(gpu_host_lib, "small_pyramid", 1, variant=1 + i) for i in range(count)]
	try:
		assert len({w.world_index() for w in gpus}) == count
		with b2.WorldGroup(gpu_host_lib, gpus) as group:
			for r in range(rounds):
				b2.step_many(ref_lib, refs, per_round)
				group.step(per_round)
				want = [w.hash() for w in refs]
				got = [w.hash() for w in gpus]
				assert got == want, f"round {r}: worlds {[i for i in range(count) if got[i] != want[i]][:8]} diverged"
			assert len(set(want)) > count // 2, "the worlds of the batch are supposed to differ"
			last = gpu_host_lib.b2GpuSeam_GetLastResult(gpus[0].world_index()).contents
			assert last.kernelLaunches >= 1 and last.gridBarriers == 0
		# out of the group again: every world back on a solver of its own
		for w, g in zip(refs[:4], gpus[:4]):
			w.step(5)
			g.step(5)
			assert g.hash() == w.hash()
	finally:
		for w in refs + gpus:
			w.destroy()
