"""Kernel-level parity through the C-ABI: feed the captured inputs of one reference solver step to
b2GpuSolverStep and compare every output with what the reference's CPU solver produced -- bit for bit
(states, manifold impulses, joint sims incl. accumulated impulses, event bit sets)."""
import numpy as np
import pytest

import box2d_b200 as b2

pytestmark = pytest.mark.gpu


def _names(files):
	return [f.name.replace(".b2cap.gz", "") for f in files]


@pytest.fixture(scope="module")
def solver():
	with b2.GpuSolver() as s:
		yield s


def _check(cap, bufs, result):
	assert np.array_equal(bufs["states"].view(np.uint32), cap.states_out.view(np.uint32)), "body states differ"
	for i, (got, want) in enumerate(zip(bufs["contacts"], cap.contacts_out)):
		if got.size:
			g = b2.contact_output_view(got).view(np.uint32)
			w = b2.contact_output_view(want).view(np.uint32)
			assert np.array_equal(g, w), f"manifold impulses differ in colour slot {i}"
			# nothing but the solver outputs may change in the contact sims
			assert np.array_equal(got, want), f"contact sims differ outside the impulses in colour slot {i}"
	for i, (got, want) in enumerate(zip(bufs["joints"], cap.joints_out)):
		assert np.array_equal(got, want), f"joint sims differ in colour slot {i}"
	assert np.array_equal(bufs["hit"][: cap.hit_bits.size], cap.hit_bits), "hit event bits differ"
	assert np.array_equal(bufs["joint"][: cap.joint_bits.size], cap.joint_bits), "joint event bits differ"
	assert bool(result.hasHitEvents) == bool(cap.has_hit_events)


# (mode, islands): persistent kernel with the island hint (island-local kernel when the bins fit), persistent kernel
# without the hint (grid-barrier kernel), one launch per stage
@pytest.mark.parametrize("mode,islands", [(0, True), (0, False), (1, False)])
def test_captured_steps_bit_exact(solver, capture_files, mode, islands):
	assert capture_files, "no golden captures"
	solver.set_mode(mode)
	island_steps = 0
	for path in capture_files:
		cap = b2.Capture(path)
		desc, result, bufs = cap.make_call(islands=islands)
		solver.step(desc, result)
		_check(cap, bufs, result)
		assert result.kernelLaunches >= 1
		if mode == 0 and result.gridBarriers == 0 and (cap.contact_count + cap.joint_count) > 0:
			island_steps += 1
	if mode == 0 and islands:
		assert island_steps > 0, "the island-local kernel never ran"
	if not islands:
		assert island_steps == 0


def test_captured_steps_bit_exact_with_island_sizes(solver, capture_files):
	"""b2GpuStepDesc::islandSizes: the bins are packed by the islands' real sizes instead of an estimate."""
	solver.set_mode(0)
	island_steps = 0
	for path in capture_files:
		cap = b2.Capture(path)
		desc, result, bufs = cap.make_call(sizes=True)
		solver.step(desc, result)
		_check(cap, bufs, result)
		island_steps += 1 if result.gridBarriers == 0 and (cap.contact_count + cap.joint_count) > 0 else 0
	assert island_steps > 0, "the island-local kernel never ran"


@pytest.mark.parametrize("blocks", [2, 4, 16])
def test_captured_steps_bit_exact_on_clusters(capture_files, blocks, monkeypatch):
	"""Force the planner to share every bin between `blocks` thread blocks of a cluster (bodies in distributed shared
	memory, colours dealt out over the blocks): same bits."""
	monkeypatch.setenv("B2GPU_CLUSTER_FORCE", str(blocks))
	cluster_steps = 0
	with b2.GpuSolver() as solver:
		for path in capture_files:
			cap = b2.Capture(path)
			desc, result, bufs = cap.make_call(islands=True)
			solver.step(desc, result)
			_check(cap, bufs, result)
			bins, per_bin = solver.island_plan()
			if bins > 0 and result.gridBarriers == 0:
				assert per_bin >= blocks
				cluster_steps += 1
	assert cluster_steps > 0, "the cluster kernel never ran"


def test_cluster_with_joint_records_left_in_global_memory(capture_files, monkeypatch):
	"""B2GPU_SPILL_JOINTS: the joint records stay in the global working copy, the bodies in distributed
	shared memory (an option the planner only takes when nothing else fits): same bits."""
	monkeypatch.setenv("B2GPU_SPILL_JOINTS", "2")  # 2 = take that plan first whenever the step has joints
	spilled = 0
	with b2.GpuSolver() as solver:
		for path in capture_files:
			cap = b2.Capture(path)
			desc, result, bufs = cap.make_call(islands=True)
			solver.step(desc, result)
			_check(cap, bufs, result)
			if cap.joint_count > 0 and solver.island_plan()[1] > 1 and result.gridBarriers == 0:
				spilled += 1
	assert spilled > 0


def test_island_failure_reruns_on_grid_kernel(capture_files, monkeypatch):
	"""When a bin does not fit its block the island kernels give up and the step is run again on the grid-barrier
	kernel from the untouched inputs: same bits, one more launch."""
	monkeypatch.setenv("B2GPU_TEST_TIGHT_BINS", "1")
	reruns = 0
	with b2.GpuSolver() as solver:
		for path in capture_files:
			cap = b2.Capture(path)
			desc, result, bufs = cap.make_call(islands=True)
			solver.step(desc, result)
			_check(cap, bufs, result)
			if solver.island_plan()[0] > 0 and result.gridBarriers > 0:
				reruns += 1
				assert result.kernelLaunches >= 3
			# and through the split-phase API (fresh buffers: the step above wrote its results into the inputs)
			desc, result, bufs = cap.make_call(islands=True)
			solver.upload(desc)
			solver.run(result)
			solver.download(desc, result)
			_check(cap, bufs, result)
	assert reruns > 0, "the rerun path never ran"


@pytest.mark.parametrize("pipelined", [True, False])
def test_phased_entry_points_with_host_threads(solver, capture_files, pipelined):
	"""The phases the seam drives from the world's workers: BeginStep, PackWork / UnpackWork (or PackRange / Wait /
	UnpackRange) called concurrently by several host threads, Submit, EndStep."""
	solver.set_mode(0)
	for path in capture_files:
		cap = b2.Capture(path)
		desc, result, bufs = cap.make_call(islands=True)
		solver.step_phased(desc, result, workers=4, pipelined=pipelined)
		_check(cap, bufs, result)
		assert result.kernelLaunches >= 1


def test_split_phase_is_repeatable(solver, capture_files):
	"""Upload once, Run twice (inputs stay pristine on the device), Download: same bits as the one-shot step."""
	solver.set_mode(0)
	cap = b2.Capture([f for f in capture_files if "falling_hinges_120" in f.name][0])
	desc, result, bufs = cap.make_call()
	solver.upload(desc)
	solver.run(result)
	solver.run(result)
	solver.download(desc, result)
	_check(cap, bufs, result)
	assert result.kernelLaunches >= 1


def test_empty_and_tiny_steps(solver):
	"""Edge cases: a world with bodies but no constraints, and zero sub-steps."""
	cap = b2.Capture(b2.ROOT / "tests" / "golden" / "contact_zoo_002.b2cap.gz")
	assert cap.contact_count == 0 and cap.joint_count == 0
	desc, result, bufs = cap.make_call()
	solver.step(desc, result)
	_check(cap, bufs, result)


@pytest.mark.parametrize("distortion", ["zero", "half", "tenfold"])
def test_wrong_island_sizes_only_cost_time(solver, capture_files, distortion):
	"""b2GpuStepDesc::islandSizes is a sizing hint: counts that are off may make a bin overflow (the device notices, the
	step is rerun on the grid-barrier kernel) or waste bins, but never change a bit of the result."""
	import ctypes

	solver.set_mode(0)
	for path in capture_files:
		cap = b2.Capture(path)
		desc, result, bufs = cap.make_call(sizes=True)
		if "sizes" not in bufs:
			continue
		sizes = np.frombuffer(bufs["sizes"], dtype=np.int32).reshape(-1, 4)
		if distortion == "zero":
			sizes[:, 0:3] = 0
		elif distortion == "half":
			sizes[:, 0:3] //= 2
		else:
			sizes[:, 0:3] *= 10
		assert desc.islandSizes == ctypes.addressof(bufs["sizes"])
		solver.step(desc, result)
		_check(cap, bufs, result)


def test_wrong_island_labels_never_change_the_result(solver, capture_files):
	"""b2GpuStepDesc::bodyIsland is a hint: labels that put the two bodies of a constraint into different islands (with or
	without sizes) are noticed on the device and the step is solved by the grid-barrier kernel instead -- same bits."""
	solver.set_mode(0)
	rng = np.random.default_rng(5)
	fallbacks = 0
	for path in capture_files:
		cap = b2.Capture(path)
		if cap.desc.islandCount < 2 or cap.contact_count + cap.joint_count == 0:
			continue
		for sizes in (False, True):
			desc, result, bufs = cap.make_call(sizes=sizes)
			labels = bufs["islands"]
			labels[:] = rng.integers(0, cap.desc.islandCount, size=labels.size, dtype=np.int32)
			solver.step(desc, result)
			_check(cap, bufs, result)
			fallbacks += 1 if result.gridBarriers > 0 else 0
	assert fallbacks > 0, "scrambled labels never reached the device check"
