#!/usr/bin/env python3
"""bench.py -- the Soft Step solver benchmark (BASELINE.json: solver ms/step and body-steps/sec per scene).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 solver
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU solver on the host cores

Workload (config.workload), default: BASELINE.json configs[1], the reference's `many_pyramids` benchmark scene
(shared/benchmarks.c:157-195: 22 000 boxes, ~58 000 touching contacts, dt = 1/60, 4 sub-steps), created with the
reference's own scene builder inside the host library.  At N > 1 every rank steps its own copy of the world on its
own GPU (independent worlds, no data-path collective, weak scaling); NCCL only reduces the result.
`--workload batch` is BASELINE.json configs[4]: 8192 independent base-10 pyramid worlds solved by
b2GpuSolverStepBatch, sharded over the ranks (strong scaling).

A "step" is one pass of the hot path (everything b2SolverTask runs, reference src/solver.c:1560-1616) over one
world step's constraints.
  value       body-steps/s with the step's inputs already resident in HBM: K x b2GpuSolverRun on one captured world
              step, device time from CUDA events on the solver's stream, L2 flushed between iterations.
  e2e         the same metric through the reference-facing call: K real b2World_Step calls of the GPU host library,
              summing b2Profile.constraints (host wall clock of the seam: joint prepare, island labels, wire packing,
              H2D, the kernels, D2H, impulse write-back).
  roofline    the step's kernels against the measured HBM peak, algorithmic bytes per SURVEY.md section 8d / DESIGN.md.
  cpu_baseline  the untouched reference (oracle/_ref, gcc -O3 build of /root/reference) on the box's host cores over a
              bounded sample of the same workload: sum of b2Profile.constraints, best worker count.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import box2d_b200 as b2  # noqa: E402

SCENES = ("many_pyramids", "large_pyramid", "joint_grid", "rain", "tumbler", "small_pyramid")
SLEEPING_SCENES = ("rain", "tumbler")  # the awake set changes at the end of a step: no re-run of a captured step
# scenes that need time to reach their steady state (rain: ~10 300 bodies in ragdolls on the ground; tumbler: everything in
# one island): extra untimed steps before the warm-up, on both arms
SETTLE_STEPS = {"rain": 400, "tumbler": 130}
METRIC = "solver_body_steps_per_sec"
UNIT = "body-steps/s"
BATCH_WORLDS = 8192


def algorithmic_bytes(bodies: int, contacts: int, joints: int, substeps: int) -> int:
	"""per coloured contact 372 + 684*s, per awake body 144*s, per coloured joint 1084*s bytes."""
	return contacts * (372 + 684 * substeps) + bodies * 144 * substeps + joints * 1084 * substeps


def measured_peaks() -> tuple[float, str]:
	path = ROOT / "MEASURED_PEAKS.json"
	if path.is_file():
		try:
			return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
		except Exception:
			pass
	return 6650.0, "fallback (B200_PROFILING.md)"


def shard_range(total: int, rank: int, world_size: int) -> tuple[int, int]:
	"""Contiguous block of units for a rank: independent worlds shard with no exchange."""
	base, extra = divmod(total, world_size)
	begin = rank * base + min(rank, extra)
	return begin, begin + base + (1 if rank < extra else 0)


def reduce_over_ranks(seconds: list[float], work: float, device=None):
	"""max over ranks of every timing, sum of the work: the only collective of the benchmark (result reduction)."""
	import torch
	import torch.distributed as dist

	t = torch.tensor(seconds, dtype=torch.float64, device=device)
	w = torch.tensor([work], dtype=torch.float64, device=device)
	if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		dist.all_reduce(w, op=dist.ReduceOp.SUM)
	return [float(x) for x in t], float(w[0])


class ClockSampler:
	"""nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

	QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
			 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

	def __init__(self, device: int):
		self.device = device
		self.samples = []
		self.proc = None

	def start(self):
		try:
			self.proc = subprocess.Popen(
				["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
				stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
			threading.Thread(target=self._read, daemon=True).start()
		except OSError:
			self.proc = None

	def _read(self):
		for line in self.proc.stdout:
			self.samples.append(line.strip())

	def stop(self) -> dict:
		if self.proc is None:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		time.sleep(0.15)
		self.proc.terminate()
		sm, mx, reasons = [], [], set()
		names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
		for line in self.samples:
			parts = [p.strip() for p in line.split(",")]
			if len(parts) < 6:
				continue
			try:
				sm.append(float(parts[0]))
				mx.append(float(parts[1]))
			except ValueError:
				continue
			for name, flag in zip(names, parts[2:6]):
				if flag.lower().startswith("active"):
					reasons.add(name)
		sm.sort()
		return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
				"reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference's CPU solver (oracle/_ref) -------------------------------------------------------------------
def load_reference():
	path = ROOT / "oracle" / "_ref" / "libbox2d_ref.so"
	if not path.is_file():
		raise RuntimeError(f"{path} missing: run python oracle/build_ref.py where /root/reference is present")
	return b2._bind_harness(ctypes.CDLL(str(path)))


def load_reference_avx2():
	"""The reference built with its optional 8-wide path (BOX2D_AVX2), or None when absent / not supported by this host."""
	path = ROOT / "oracle" / "_ref" / "libbox2d_ref_avx2.so"
	try:
		flags = open("/proc/cpuinfo").read()
	except OSError:
		flags = ""
	if not path.is_file() or " avx2" not in flags:
		return None
	return b2._bind_harness(ctypes.CDLL(str(path)))


def scene_config(scene: str, bodies: int, contacts: int, joints: int, substeps: int = 4) -> dict:
	"""The workload as both arms name it (identical keys and values, so that the two lines can be compared field by field)."""
	return {"workload": scene, "settle_steps": SETTLE_STEPS.get(scene, 0), "bodies": bodies, "contacts": contacts, "joints": joints,
			"substeps": substeps, "dt": 1.0 / 60.0}


def cpu_solver_run(scene: str, workers: int, warmup: int, steps: int, lib=None, replicas: int = 1) -> dict:
	"""Time the reference's CPU solver: sum of b2Profile.constraints over `steps` world steps after `warmup`.  replicas > 1:
	that many independent copies of the scene stepped concurrently, one host thread each driving `workers` workers (the
	multi-GPU arm steps one world per GPU); the times are the slowest replica's."""
	lib = lib or load_reference()
	worlds = [b2.World(lib, scene, workers) for _ in range(replicas)]
	try:
		def warm(w):
			w.step(SETTLE_STEPS.get(scene, 0) + warmup)

		def timed(w):
			return w.bench(steps)

		with ThreadPoolExecutor(max_workers=replicas) as pool:
			list(pool.map(warm, worlds))
			bodies = worlds[0].counters()["awakeBodyCount"]
			results = list(pool.map(timed, worlds))
		c = worlds[0].counters()
	finally:
		for w in worlds:
			w.destroy()
	r = max(results, key=lambda x: x["constraints_ms"])
	r.update(bodies=bodies, contacts=sum(c["colorCounts"]) - c["jointCount"], joints=c["jointCount"], workers=workers, replicas=replicas)
	return r


def best_cpu_baseline(scene: str, warmup: int, steps: int, replicas: int = 1) -> dict:
	cores = os.cpu_count() or 1
	share = max(1, min(cores // replicas, 32))
	candidates = sorted({w for w in (1, 4, 8, 16, 24, 32, share) if w <= share})
	best = None
	tried = {}
	for workers in candidates:
		r = cpu_solver_run(scene, workers, warmup, steps, replicas=replicas)
		tried[workers] = r["constraints_ms"] / steps
		if best is None or r["constraints_ms"] < best["constraints_ms"]:
			best = r
	best["tried_ms_per_step"] = tried
	best["host_cores"] = cores
	best["build"] = "default (SSE2, 4 lanes)"
	best["sse2_ms_per_step"] = best["constraints_ms"] / steps
	# the reference's optional AVX2 build (8 lanes) at the best worker count: the number to beat is the best CPU configuration
	# (BASELINE.md section 3), so when it is faster it becomes the baseline
	avx2 = load_reference_avx2()
	best["avx2_ms_per_step"] = None
	if avx2:
		r = cpu_solver_run(scene, best["workers"], warmup, steps, avx2, replicas=replicas)
		best["avx2_ms_per_step"] = r["constraints_ms"] / steps
		if r["constraints_ms"] < best["constraints_ms"]:
			keep = {k: best[k] for k in ("tried_ms_per_step", "host_cores", "sse2_ms_per_step", "avx2_ms_per_step")}
			best = dict(r, **keep)
			best["build"] = "BOX2D_AVX2 (8 lanes)"
	return best


def cpu_batch_baseline(worlds: int, warmup: int, steps: int) -> dict:
	"""The batch on the host: ALL `worlds` independent worlds, stepped concurrently by a thread pool with one worker per world
	(worlds are independent, include/box2d/box2d.h:31-32) in chunks that fit B2_MAX_WORLDS = 128
	(include/box2d/constants.h:38-40).  Solver time of a batch step = sum(b2Profile.constraints) / threads, i.e. the
	solver-only share of a perfectly parallel host loop; the wall clock of the whole loop is reported beside it."""
	lib = load_reference()
	threads = min(os.cpu_count() or 1, 32)
	chunk = max(threads, (112 // threads) * threads)
	constraints_ms = 0.0
	step_ms = 0.0
	wall = 0.0
	bodies = 0
	done = 0
	with ThreadPoolExecutor(max_workers=threads) as pool:
		while done < worlds:
			n = min(chunk, worlds - done)
			ws = [b2.World(lib, "small_pyramid", 1) for _ in range(n)]
			groups = [ws[i::threads] for i in range(threads)]

			def warm(group):
				for w in group:
					w.step(warmup)

			def run(group):
				c = s = 0.0
				for w in group:
					r = w.bench(steps)
					c += r["constraints_ms"]
					s += r["step_ms"]
				return c, s

			list(pool.map(warm, groups))
			bodies = ws[0].counters()["awakeBodyCount"]
			t0 = time.perf_counter()
			for c, s in pool.map(run, groups):
				constraints_ms += c
				step_ms += s
			wall += time.perf_counter() - t0
			for w in ws:
				w.destroy()
			done += n
	solver_s = constraints_ms * 1e-3 / threads
	return {"value": worlds * steps * bodies / solver_s, "unit": UNIT, "cores": os.cpu_count() or 1, "threads": threads, "kind": "reference",
			"us_per_world_step": constraints_ms * 1e3 / (worlds * steps),
			"solver_ms_per_batch_step": solver_s * 1e3 / steps,
			"whole_step_ms_per_batch_step": step_ms / threads / steps,
			"wall_ms_per_batch_step": wall * 1e3 / steps,
			"sample": f"all {worlds} base-10 pyramid worlds x {steps} steps after {warmup} warm-up steps, in chunks of {chunk} "
					  f"(B2_MAX_WORLDS), one worker per world, {threads} host threads; sum of b2Profile.constraints / threads"}


def run_reference_arm(args) -> int:
	"""--impl reference: the reference's own CPU implementation of the path, all host threads it can use."""
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return 0
	if args.workload == "batch":
		b = cpu_batch_baseline(BATCH_WORLDS, args.warmup, args.steps)
		line = {"impl": "reference", "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
				"warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
				"dtype": "f32", "data": "synthetic", "config": {"workload": f"batch of {BATCH_WORLDS} small_pyramid worlds"},
				"cpu_baseline": b, "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
				"gpu_launches": 0}
		print(json.dumps(line))
		return 0
	# the GPU arm at N GPUs steps N independent copies of the world (weak scaling): so does this arm, on the host's cores
	replicas = max(1, args.gpus)
	best = best_cpu_baseline(args.workload, args.warmup, args.steps, replicas)
	ms = best["constraints_ms"] / args.steps
	value = replicas * best["bodies"] * args.steps / (best["constraints_ms"] * 1e-3)
	line = {
		"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
		"warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
		"dtype": "f32", "data": "synthetic",
		"config": scene_config(args.workload, best["bodies"], best["contacts"], best["joints"]),
		"details": {"timed": "sum of b2Profile.constraints (reference src/solver.c:1561,1615), slowest of the concurrently stepped worlds",
					"worlds": replicas, "workers_per_world": best["workers"], "build": best["build"],
					"whole_step_ms": best["step_ms"] / args.steps},
		"cpu_baseline": {"value": value, "unit": UNIT, "cores": best["host_cores"], "best_workers": best["workers"], "kind": "reference",
						 "build": best["build"],
						 "sample": f"{args.steps} steps of {replicas} x {args.workload} after {args.warmup} warm-up steps; best of worker counts "
								   f"{sorted(best['tried_ms_per_step'])} per world and of the default / AVX2 builds on {best['host_cores']} host cores",
						 "ms_per_step_by_workers": best["tried_ms_per_step"],
						 "sse2_build_ms_per_step": best["sse2_ms_per_step"],
						 "avx2_build_ms_per_step": best["avx2_ms_per_step"]},
		"e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}
	if args.batch:
		line["batch"] = cpu_batch_baseline(BATCH_WORLDS, 30, min(args.steps, 10))
	print(json.dumps(line))
	return 0


# ---- the B200 arm ----------------------------------------------------------------------------------------------------
def _setup_distributed():
	import torch
	import torch.distributed as dist

	world_size = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))
	if not torch.cuda.is_available():
		raise RuntimeError("bench.py needs a CUDA device: the solver has no CPU fallback")
	torch.cuda.set_device(local_rank)
	os.environ["B2GPU_DEVICE"] = str(local_rank)
	# the library's own host threads (one-call entry points): the ranks of one box share its cores
	os.environ.setdefault("B2GPU_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(1, world_size))))
	if world_size > 1:
		# The ranks of one box share its cores: every rank keeps to its own share of them (set before NCCL and the world's
		# workers start their threads, which inherit it).  The workers spin across the short gaps between the host passes of
		# a step; two ranks meeting on one core wait for each other's time slices.  BENCH_PIN_RANKS=0: leave it to the OS.
		allowed = sorted(os.sched_getaffinity(0))
		share = len(allowed) // world_size
		if share >= 2 and os.environ.get("BENCH_PIN_RANKS", "1") != "0":
			os.sched_setaffinity(0, allowed[local_rank * share:(local_rank + 1) * share])
		dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
	return world_size, rank, local_rank


def measured_traffic(workload: str):
	"""DRAM bytes per step of the workload's kernels from the committed ncu --set full captures (profiles/), or None."""
	for name in ("r02_traffic.json", "r01_traffic.json"):
		try:
			entry = json.loads((ROOT / "profiles" / name).read_text()).get(workload)
		except (OSError, ValueError):
			continue
		if entry:
			return entry["bytes"]
	return None


def run_scene(args) -> int:
	import torch
	import torch.distributed as dist

	world_size, rank, local_rank = _setup_distributed()
	distributed = world_size > 1
	host = b2.host_lib()
	host.b2GpuSeam_InstallPinnedAllocator()
	flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

	def sync_all():
		torch.cuda.synchronize()
		if distributed:
			dist.barrier()
		torch.cuda.synchronize()

	sampler = ClockSampler(local_rank)
	# host phases of the GPU arm (collide, pack/unpack, finalize): the ranks of one box share its cores
	# (with several ranks one core per rank is left to the process's other threads: the helpers spin across the short gaps
	# between the host passes, an oversubscribed box would make them wait for each other's time slices)
	cores = os.cpu_count() or 1
	workers = max(1, min(cores, 16)) if world_size == 1 else max(1, min(cores // world_size - 1, 16))
	scene = args.workload
	with b2.World(host, scene, workers) as world:
		world.step(SETTLE_STEPS.get(scene, 0) + args.warmup)
		widx = world.world_index()
		counters = world.counters()
		bodies = counters["awakeBodyCount"]
		joints = counters["jointCount"]
		contacts = sum(counters["colorCounts"]) - joints

		# ---- e2e: K real b2World_Step calls, host buffers, copies inside the timed region ----
		totals = b2.SeamTotals()
		host.b2GpuSeam_GetTotals(widx, ctypes.byref(totals), 1)
		sync_all()
		sampler.start()
		t0 = time.perf_counter()
		e2e = world.bench(args.steps)
		e2e_wall = time.perf_counter() - t0
		host.b2GpuSeam_GetTotals(widx, ctypes.byref(totals), 1)
		last = host.b2GpuSeam_GetLastResult(widx).contents
		e2e_split = {"pack_ms": float(last.uploadMs), "h2d_kernels_d2h_ms": float(last.waitMs), "unpack_ms": float(last.scatterMs),
					 "abi_total_ms": float(last.totalMs), "kernel_ms": float(last.kernelMs)}
		e2e_launches = int(totals.launches)
		h2d = totals.h2dBytes / max(1, totals.steps)
		d2h = totals.d2hBytes / max(1, totals.steps)
		e2e_kernel_s = totals.kernelMs * 1e-3
		stage_ms = [totals.stageMs[i] / max(1, totals.steps) for i in range(8)]
		grid_barriers = totals.gridBarriers / max(1, totals.steps)
		island_plan = None
		launches_per_step = totals.launches / max(1, totals.steps)

		# ---- value: the captured step resident in HBM, K x Run, CUDA events, L2 flushed in between ----
		resident = scene not in SLEEPING_SCENES
		resident_stats = None
		abi_split = {"seam_ms": totals.seamMs / max(1, totals.steps), "pack_ms": totals.packMs / max(1, totals.steps),
					 "h2d_kernels_d2h_ms": totals.waitMs / max(1, totals.steps), "unpack_ms": totals.unpackMs / max(1, totals.steps),
					 "before_begin_ms": totals.beforeMs / max(1, totals.steps)}
		kernel_s = e2e_kernel_s
		launches = 0
		substeps = 4
		if resident:
			# The desc still points at the world's arrays.  After b2World_Step returned they hold what the NEXT step
			# would start from (finalize reset the state deltas, src/solver.c:611-612; manifolds carry the stored
			# impulses), i.e. a valid, representative solver input with the same constraint graph.
			host.b2GpuSeam_FlushImpulses(widx)  # (deferred impulses: the manifolds receive what the last step computed)
			desc = host.b2GpuSeam_GetLastDesc(widx).contents
			substeps = int(desc.subStepCount)
			with b2.GpuSolver(device=local_rank) as solver:
				# Steady state of the resident mode: one whole step first, so that the device holds the world's
				# contacts and bodies; then the host does to the outputs what b2FinalizeBodiesTask does (deltas reset, transient
				# flags cleared, src/solver.c:611-612, :632) and the next step's inputs are uploaded -- light records only, as
				# in the e2e steps above -- and stay resident for the timed runs.
				result = b2.StepResult()
				solver.step(desc, result)
				n = int(desc.awakeBodyCount)
				st = np.ctypeslib.as_array(ctypes.cast(desc.states, ctypes.POINTER(ctypes.c_uint32)), shape=(n, 8))
				st[:, 4:8] = np.array([0.0, 0.0, 1.0, 0.0], dtype=np.float32).view(np.uint32)
				st[:, 3] &= np.uint32(0xFFFFFF97)
				solver.upload(desc)
				resident_stats = solver.resident_stats()
				for _ in range(max(3, args.warmup)):
					solver.run(result)
				sync_all()
				kernel_ms = []
				stage_ms = [0.0] * 8
				for _ in range(args.steps):
					flush.zero_()
					torch.cuda.synchronize()
					solver.run(result)
					kernel_ms.append(float(result.kernelMs))
					for i in range(8):
						stage_ms[i] += float(result.stageMs[i]) / args.steps
				launches = int(result.kernelLaunches) * args.steps
				launches_per_step = int(result.kernelLaunches)
				grid_barriers = int(result.gridBarriers)
				island_plan = list(solver.island_plan())
			kernel_s = sum(kernel_ms) * 1e-3
		sync_all()
		clocks = sampler.stop()

	(kernel_s, e2e_s), total_work = reduce_over_ranks([kernel_s, e2e["constraints_ms"] * 1e-3], float(bodies * args.steps), "cuda")

	if rank == 0:
		ms_per_step = kernel_s * 1e3 / args.steps
		peak, peak_source = measured_peaks()
		alg = algorithmic_bytes(bodies, contacts, joints, substeps)
		achieved = alg / (ms_per_step * 1e-3) / 1e9
		cpu = None
		if args.cpu_baseline and not distributed:
			sample_steps = min(args.steps, 60)
			b = best_cpu_baseline(scene, args.warmup, sample_steps)
			cpu = {"value": b["bodies"] * sample_steps / (b["constraints_ms"] * 1e-3), "unit": UNIT, "cores": b["host_cores"],
				   "best_workers": b["workers"], "build": b["build"],
				   "kind": "reference", "ms_per_step": b["constraints_ms"] / sample_steps, "whole_step_ms": b["step_ms"] / sample_steps,
				   "sample": f"{sample_steps} steps of {scene} after {args.warmup} warm-up steps, sum of "
							 f"b2Profile.constraints; best of worker counts {sorted(b['tried_ms_per_step'])} and of the default / AVX2 "
							 f"builds on {b['host_cores']} host cores (oracle/_ref = untouched reference, gcc -O3)",
				   "ms_per_step_by_workers": b["tried_ms_per_step"],
				   "sse2_build_ms_per_step": b["sse2_ms_per_step"],
				   "avx2_build_ms_per_step": b["avx2_ms_per_step"]}
		line = {
			"metric": METRIC, "value": total_work / kernel_s, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
			"warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
			"vs_baseline": None, "dtype": "f32", "data": "synthetic",
			"config": scene_config(scene, bodies, contacts, joints, substeps),
			"details": {"colors": sum(1 for c in counters["colorCounts"][:23] if c > 0),
						"parallelism": f"{world_size} independent world(s), one per GPU", "host_workers_per_world": workers,
						"rank_cores": len(os.sched_getaffinity(0)),
						"l2": "flushed (256 MiB write) between timed iterations" if resident else
							  "not flushed: kernels timed inside the real steps (awake set changes every step)",
						"timed": "CUDA events on the solver stream around the step's kernels, inputs resident"},
			"e2e": {"value": total_work / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3 / args.steps,
					"h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
					"timed": "sum of b2Profile.constraints over K b2World_Step calls of libbox2d_b200.so",
					"whole_step_ms": e2e["step_ms"] / args.steps, "wall_ms_per_step": e2e_wall * 1e3 / args.steps,
					"kernel_ms_per_step": e2e_kernel_s * 1e3 / args.steps, "mean_split": abi_split, "last_step_split": e2e_split},
			"gpu_launches": launches + e2e_launches,
			"roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
						 "traffic": measured_traffic(scene), "algorithmic_bytes_per_launch": alg, "peak_source": peak_source,
						 "kernels_per_step": launches_per_step, "grid_barriers_per_step": grid_barriers,
						 "island_bins_blocks_per_bin": island_plan,
						 "resident_upload": None if resident_stats is None else {"full_contacts": resident_stats[0], "dirty_bodies": resident_stats[1]},
						 "note": "all kernels of the step (scatter or partition kernel + island or cluster kernel, or the grid-barrier kernel); see DESIGN.md"},
			"stage_ms_per_step": {n: stage_ms[i] for i, n in enumerate(b2.STAGE_NAMES)},
			"clocks": clocks,
		}
		if cpu is not None:
			line["cpu_baseline"] = cpu
	# north_star's multi-GPU configuration rides along in the same line: the 8192-world batch sharded over the N ranks
	batch = measure_batch(args, world_size, rank, local_rank, args.cpu_baseline and not distributed) if args.batch else None
	if rank == 0:
		if batch is not None:
			line["batch"] = batch
		print(json.dumps(line))
	if distributed:
		dist.destroy_process_group()
	return 0


def build_batch(host, worlds: int, warmup: int):
	"""`worlds` independent copies of one settled base-10 pyramid world: every world gets its own host arrays."""
	with b2.World(host, "small_pyramid", 1) as w:
		w.step(warmup)
		desc = host.b2GpuSeam_GetLastDesc(w.world_index()).contents
		n = desc.awakeBodyCount

		def grab(ptr, nbytes):
			if nbytes == 0:
				return np.zeros(0, np.uint8)
			return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,)).copy()

		states = grab(desc.states, n * b2.STATE_SIZE)
		sims = grab(desc.sims, n * b2.SIM_SIZE)
		labels = np.ctypeslib.as_array(ctypes.cast(desc.bodyIsland, ctypes.POINTER(ctypes.c_int)), shape=(n,)).copy()
		sizes = grab(desc.islandSizes, desc.islandCount * ctypes.sizeof(b2.IslandSize)) if desc.islandSizes else None
		colors = [(grab(desc.colors[c].contactSims, desc.colors[c].contactCount * b2.CONTACT_SIZE), desc.colors[c].contactCount)
				  for c in range(desc.activeColorCount)]
		template = b2.StepDesc.from_buffer_copy(bytes(desc))
		contacts = sum(cnt for _, cnt in colors)

	descs = (b2.StepDesc * worlds)()
	results = (b2.StepResult * worlds)()
	keep = []
	for i in range(worlds):
		d = b2.StepDesc.from_buffer_copy(bytes(template))
		st, sm, lb = states.copy(), sims.copy(), labels.copy()
		cs = [a.copy() for a, _ in colors]
		d.states, d.sims, d.bodyIsland = st.ctypes.data, sm.ctypes.data, lb.ctypes.data
		d.islandSizes = sizes.ctypes.data if sizes is not None else None  # read-only, shared by the copies
		for c, arr in enumerate(cs):
			d.colors[c].contactSims = arr.ctypes.data
		ctypes.memmove(ctypes.byref(descs[i]), ctypes.byref(d), ctypes.sizeof(b2.StepDesc))
		keep.append((st, sm, lb, cs))
	descs._island_sizes = sizes  # keep-alive
	pristine = (states, [a for a, _ in colors])
	return descs, results, keep, pristine, n, contacts, int(template.subStepCount)


def measure_batch(args, world_size: int, rank: int, local_rank: int, cpu_baseline: bool):
	"""BASELINE.json configs[4]: 8192 independent base-10 pyramid worlds sharded over the ranks (contiguous blocks, no
	exchange), solved by b2GpuSolverStepBatch.  Returns the record on rank 0, None elsewhere."""
	import torch
	import torch.distributed as dist

	distributed = world_size > 1
	# the library packs / unpacks a batch on its own threads: the ranks of one box share its cores
	host = b2.host_lib()
	begin, end = shard_range(BATCH_WORLDS, rank, world_size)
	worlds = end - begin
	descs, results, keep, pristine, bodies, contacts, substeps = build_batch(host, worlds, max(args.warmup, 30))
	flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

	def sync_all():
		torch.cuda.synchronize()
		if distributed:
			dist.barrier()
		torch.cuda.synchronize()

	def restore():
		for st, _, _, cs in keep:
			st[:] = pristine[0]
			for a, p in zip(cs, pristine[1]):
				a[:] = p

	steps = min(args.steps, 50)
	sampler = ClockSampler(local_rank)
	with b2.GpuSolver(device=local_rank) as solver:
		# value: inputs resident, K x RunBatch
		solver.upload_batch(descs)
		r = b2.StepResult()
		for _ in range(3):
			solver.run_batch(r)
		sync_all()
		sampler.start()
		kernel_ms = []
		torch.cuda.profiler.start()  # lets `ncu --profile-from-start off` skip the thousands of one-world warm-up launches
		for _ in range(steps):
			flush.zero_()
			torch.cuda.synchronize()
			solver.run_batch(r)
			kernel_ms.append(float(r.kernelMs))
		torch.cuda.profiler.stop()
		launches = int(r.kernelLaunches) * steps
		grid_barriers = int(r.gridBarriers)
		island_plan = list(solver.island_plan())
		# e2e: the whole b2GpuSolverStepBatch from host arrays, inputs restored untimed
		e2e_steps = max(3, min(steps, 10))
		e2e_s = 0.0
		for _ in range(e2e_steps):
			restore()
			sync_all()
			t0 = time.perf_counter()
			solver.step_batch(descs, results)
			e2e_s += time.perf_counter() - t0
		h2d, d2h = int(results[0].h2dBytes), int(results[0].d2hBytes)
		split = {"pack_ms": float(results[0].uploadMs), "h2d_kernels_d2h_ms": float(results[0].waitMs),
				 "unpack_ms": float(results[0].scatterMs)}
		launches += int(results[0].kernelLaunches) * e2e_steps
	sync_all()
	clocks = sampler.stop()

	kernel_s = sum(kernel_ms) * 1e-3
	(kernel_s, e2e_per_step), total_work = reduce_over_ranks([kernel_s, e2e_s / e2e_steps], float(worlds * bodies * steps), "cuda")
	total_worlds = BATCH_WORLDS
	if rank != 0:
		return None
	ms_per_step = kernel_s * 1e3 / steps
	peak, peak_source = measured_peaks()
	alg = algorithmic_bytes(bodies, contacts, 0, substeps) * worlds
	achieved = alg / (ms_per_step * 1e-3) / 1e9
	record = {
		"metric": METRIC, "value": total_work / kernel_s, "unit": UNIT, "n_gpus": world_size, "steps": steps,
		"warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
		"vs_baseline": None, "dtype": "f32", "data": "synthetic",
		"config": {"workload": f"batch of {total_worlds} small_pyramid worlds", "worlds": total_worlds, "worlds_per_gpu": worlds,
				   "bodies_per_world": bodies, "contacts_per_world": contacts, "substeps": substeps,
				   "world_steps_per_sec": total_worlds * steps / kernel_s,
				   "parallelism": f"worlds sharded over {world_size} GPU(s), no exchange",
				   "l2": "flushed (256 MiB write) between timed iterations"},
		"e2e": {"value": total_worlds * bodies / e2e_per_step, "unit": UNIT, "ms_per_step": e2e_per_step * 1e3,
				"h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "last_step_split": split,
				"timed": "b2GpuSolverStepBatch wall clock: host packing (library threads) + H2D + kernels + D2H + write-back"},
		"gpu_launches": launches,
		"roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
					 "traffic": measured_traffic("batch") if world_size == 1 else None,
					 "algorithmic_bytes_per_launch": alg, "peak_source": peak_source, "grid_barriers_per_step": grid_barriers,
					 "island_bins_blocks_per_bin": island_plan},
		"clocks": clocks,
	}
	if cpu_baseline:
		record["cpu_baseline"] = cpu_batch_baseline(total_worlds, 30, min(steps, 10))
		record["live_worlds"] = live_group_run(host, min(worlds, 1024), 100)
	return record


def live_group_run(host, count: int, steps: int) -> dict:
	"""World-level batched step (b2GpuSeam_CreateGroup): `count` LIVE worlds -- every world its own slightly different pile,
	offset derived from its index -- stepped concurrently through b2World_Step for `steps` steps: broad phase, narrow phase
	and finalize per world on the host, one b2GpuSolverStepBatch per round.  A sample of the worlds is stepped by the
	reference as well and compared (state hash).  The wall clock covers the WHOLE step of every world."""
	ref = load_reference()
	sample = min(count, 32)
	worlds = [b2.World(host, "small_pyramid", 1, variant=1 + i) for i in range(count)]
	refs = [b2.World(ref, "small_pyramid", 1, variant=1 + i) for i in range(sample)]
	try:
		with b2.WorldGroup(host, worlds) as group:
			group.step(10)
			t0 = time.perf_counter()
			group.step(steps)
			wall = time.perf_counter() - t0
			last = host.b2GpuSeam_GetLastResult(worlds[0].world_index()).contents
			split = {"pack_ms": float(last.uploadMs), "h2d_kernels_d2h_ms": float(last.waitMs), "unpack_ms": float(last.scatterMs),
					 "kernel_ms": float(last.kernelMs)}
		b2.step_many(ref, refs, 10 + steps)
		same = all(worlds[i].hash() == refs[i].hash() for i in range(sample))
		bodies = worlds[0].counters()["awakeBodyCount"]
	finally:
		for w in worlds + refs:
			w.destroy()
	return {"worlds": count, "steps": steps, "wall_ms_per_batch_step": wall * 1e3 / steps,
			"whole_step_body_steps_per_sec": count * bodies * steps / wall, "last_batch_split": split,
			"parity": f"state hash of the first {sample} worlds equals the reference's after {10 + steps} steps: {same}",
			"parity_ok": bool(same),
			"note": "one OS thread per world meets the others at the seam; the wall clock includes every world's broad phase, narrow "
					"phase and finalize on the host"}


def run_batch(args) -> int:
	import torch.distributed as dist

	world_size, rank, local_rank = _setup_distributed()
	record = measure_batch(args, world_size, rank, local_rank, args.cpu_baseline and world_size == 1)
	if record is not None:
		print(json.dumps(record))
	if world_size > 1:
		dist.destroy_process_group()
	return 0


def main() -> int:
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=100)
	ap.add_argument("--warmup", type=int, default=20)
	ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
	ap.add_argument("--workload", "--scene", dest="workload", default="many_pyramids", choices=SCENES + ("batch",))
	ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
	ap.add_argument("--no-batch", dest="batch", action="store_false",
					help="scene workloads: skip the 8192-world batch sub-record (north_star's multi-GPU configuration)")
	args = ap.parse_args()
	args.warmup = max(3, args.warmup)
	if args.impl == "reference":
		return run_reference_arm(args)
	if args.workload == "batch":
		return run_batch(args)
	return run_scene(args)


if __name__ == "__main__":
	sys.exit(main())
