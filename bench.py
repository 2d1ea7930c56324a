#!/usr/bin/env python3
"""bench.py -- the Soft Step solver benchmark (BASELINE.json: solver ms/step and body-steps/sec per scene).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 solver
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU solver on the host cores

Workload (config.workload): BASELINE.json configs[1], the reference's `many_pyramids` benchmark scene
(shared/benchmarks.c:157-195: 22 000 boxes, ~58 000 touching contacts, dt = 1/60, 4 sub-steps), created with the
reference's own scene builder inside the host library.  At N > 1 every rank steps its own copy of the world on its
own GPU (independent worlds, no data-path collective, weak scaling); NCCL is used only to reduce the result.

A "step" is one pass of the hot path (everything b2SolverTask runs, reference src/solver.c:1560-1616) over one
world step's constraints.
  value       body-steps/s with the step's inputs already resident in HBM: K x b2GpuSolverRun on one captured world
              step, device time from CUDA events on the solver's stream, L2 flushed between iterations.
  e2e         the same metric through the reference-facing call: K real b2World_Step calls of the GPU host library,
              summing b2Profile.constraints (host wall clock of the seam: host joint prepare, H2D of the reference's
              own arrays from pinned memory, the step kernel, D2H of states / impulses / joints / event bits).
  roofline    the step kernel (the only kernel of the step) against the measured HBM peak, algorithmic bytes per
              SURVEY.md section 8d / DESIGN.md.
  cpu_baseline  the untouched reference (oracle/_ref, gcc -O3 build of /root/reference) on the box's host cores over a
              bounded sample of the same workload: sum of b2Profile.constraints.
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
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import box2d_b200 as b2  # noqa: E402

SCENE_STEPS = {"many_pyramids": 200, "large_pyramid": 500, "joint_grid": 500, "rain": 1000, "tumbler": 750,
			   "small_pyramid": 200}
METRIC = "solver_body_steps_per_sec"
UNIT = "body-steps/s"


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


def load_reference():
	path = ROOT / "oracle" / "_ref" / "libbox2d_ref.so"
	if not path.is_file():
		raise RuntimeError(f"{path} missing: run python oracle/build_ref.py where /root/reference is present")
	return b2._bind_harness(ctypes.CDLL(str(path)))


def cpu_solver_run(scene: str, workers: int, warmup: int, steps: int) -> dict:
	"""Time the reference's CPU solver: sum of b2Profile.constraints over `steps` world steps after `warmup`."""
	lib = load_reference()
	with b2.World(lib, scene, workers) as w:
		w.step(warmup)
		bodies = w.counters()["awakeBodyCount"]
		r = w.bench(steps)
		c = w.counters()
	r.update(bodies=bodies, contacts=sum(c["colorCounts"]) - c["jointCount"], joints=c["jointCount"], workers=workers)
	return r


def best_cpu_baseline(scene: str, warmup: int, steps: int) -> dict:
	cores = os.cpu_count() or 1
	candidates = sorted({w for w in (1, 4, 8, 16, 24, 32, min(cores, 32)) if w <= min(cores, 32)})
	best = None
	tried = {}
	for workers in candidates:
		r = cpu_solver_run(scene, workers, warmup, steps)
		tried[workers] = r["constraints_ms"] / steps
		if best is None or r["constraints_ms"] < best["constraints_ms"]:
			best = r
	best["tried_ms_per_step"] = tried
	best["host_cores"] = cores
	return best


def run_reference_arm(args) -> int:
	"""--impl reference: the reference's own CPU implementation of the path, all host threads it can use."""
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return 0
	best = best_cpu_baseline(args.scene, args.warmup, args.steps)
	ms = best["constraints_ms"] / args.steps
	value = best["bodies"] * args.steps / (best["constraints_ms"] * 1e-3)
	line = {
		"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
		"warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
		"dtype": "f32", "data": "synthetic",
		"config": {"workload": args.scene, "bodies": best["bodies"], "contacts": best["contacts"], "joints": best["joints"],
				   "substeps": 4, "dt": 1.0 / 60.0, "timed": "sum of b2Profile.constraints (reference src/solver.c:1561,1615)"},
		"cpu_baseline": {"value": value, "unit": UNIT, "cores": best["workers"], "kind": "reference",
						 "sample": f"{args.steps} steps of {args.scene} after {args.warmup} warm-up steps; best of worker counts "
								   f"{sorted(best['tried_ms_per_step'])} on {best['host_cores']} host cores",
						 "ms_per_step_by_workers": best["tried_ms_per_step"]},
		"e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}
	print(json.dumps(line))
	return 0


def run_gpu_arm(args) -> int:
	import torch
	import torch.distributed as dist

	world_size = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))
	distributed = world_size > 1
	if not torch.cuda.is_available():
		raise RuntimeError("bench.py needs a CUDA device: the solver has no CPU fallback")
	torch.cuda.set_device(local_rank)
	os.environ["B2GPU_DEVICE"] = str(local_rank)
	if distributed:
		dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

	host = b2.host_lib()
	host.b2GpuSeam_InstallPinnedAllocator()
	flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

	def sync_all():
		torch.cuda.synchronize()
		if distributed:
			dist.barrier()
		torch.cuda.synchronize()

	sampler = ClockSampler(local_rank)
	workers = min(os.cpu_count() or 1, 16)  # host phases of the GPU arm (collide, pack/unpack, finalize)
	with b2.World(host, args.scene, workers) as world:
		world.step(args.warmup)
		widx = world.world_index()
		counters = world.counters()
		bodies = counters["awakeBodyCount"]
		joints = counters["jointCount"]
		contacts = sum(counters["colorCounts"]) - joints

		# ---- e2e: K real b2World_Step calls, host buffers, copies inside the timed region ----
		sync_all()
		sampler.start()
		t0 = time.perf_counter()
		e2e = world.bench(args.steps)
		e2e_wall = time.perf_counter() - t0
		last = host.b2GpuSeam_GetLastResult(widx).contents
		h2d, d2h = int(last.h2dBytes), int(last.d2hBytes)
		e2e_launches = int(last.kernelLaunches) * args.steps
		grid_barriers = int(last.gridBarriers)
		e2e_kernel_ms_last = float(last.kernelMs)
		e2e_split = {"upload_enqueue_ms": float(last.uploadMs), "h2d_ms": float(last.h2dMs), "wait_ms": float(last.waitMs),
					 "scatter_ms": float(last.scatterMs), "abi_total_ms": float(last.totalMs)}

		# ---- value: the captured step resident in HBM, K x Run, CUDA events, L2 flushed in between ----
		desc = host.b2GpuSeam_GetLastDesc(widx).contents
		substeps = int(desc.subStepCount)
		with b2.GpuSolver(device=local_rank) as solver:
			# The desc still points at the world's arrays.  After b2World_Step returned they hold what the NEXT
			# step would start from (finalize reset the state deltas, src/solver.c:611-612; manifolds carry the stored
			# impulses), i.e. a valid, representative solver input with the same constraint graph.
			solver.upload(desc)
			result = b2.StepResult()
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
					stage_ms[i] += float(result.stageMs[i])
			launches = int(result.kernelLaunches) * args.steps
		sync_all()
		clocks = sampler.stop()

	total_kernel_s = sum(kernel_ms) * 1e-3
	times = torch.tensor([total_kernel_s, e2e["constraints_ms"] * 1e-3], dtype=torch.float64, device="cuda")
	work = torch.tensor([float(bodies * args.steps)], dtype=torch.float64, device="cuda")
	if distributed:
		dist.all_reduce(times, op=dist.ReduceOp.MAX)  # max over ranks
		dist.all_reduce(work, op=dist.ReduceOp.SUM)   # result reduction only: no data-path collective
	total_kernel_s, e2e_s = float(times[0]), float(times[1])
	total_work = float(work[0])

	if rank == 0:
		ms_per_step = total_kernel_s * 1e3 / args.steps
		peak, peak_source = measured_peaks()
		alg = algorithmic_bytes(bodies, contacts, joints, substeps)
		achieved = alg / (ms_per_step * 1e-3) / 1e9
		cpu = None
		if args.cpu_baseline and not distributed:
			sample_steps = min(args.steps, 60)
			b = best_cpu_baseline(args.scene, args.warmup, sample_steps)
			cpu = {"value": b["bodies"] * sample_steps / (b["constraints_ms"] * 1e-3), "unit": UNIT, "cores": b["workers"],
				   "kind": "reference", "ms_per_step": b["constraints_ms"] / sample_steps,
				   "sample": f"{sample_steps} steps of {args.scene} after {args.warmup} warm-up steps, sum of "
							 f"b2Profile.constraints; best of worker counts {sorted(b['tried_ms_per_step'])} on "
							 f"{b['host_cores']} host cores (oracle/_ref = untouched reference, gcc -O3 SSE2)",
				   "ms_per_step_by_workers": b["tried_ms_per_step"]}
		line = {
			"metric": METRIC, "value": total_work / total_kernel_s, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
			"warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
			"vs_baseline": None, "dtype": "f32", "data": "synthetic",
			"config": {"workload": args.scene, "bodies": bodies, "contacts": contacts, "joints": joints, "substeps": substeps,
					   "dt": 1.0 / 60.0, "colors": sum(1 for c in counters["colorCounts"][:23] if c > 0),
					   "parallelism": f"{world_size} independent world(s), one per GPU",
					   "l2": "flushed (256 MiB write) between timed iterations",
					   "timed": "CUDA events on the solver stream around the step kernel, inputs resident"},
			"e2e": {"value": total_work / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3 / args.steps,
					"h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
					"timed": "sum of b2Profile.constraints over K b2World_Step calls of libbox2d_b200.so",
					"whole_step_ms": e2e["step_ms"] / args.steps, "wall_ms_per_step": e2e_wall * 1e3 / args.steps,
					"kernel_ms_last_step": e2e_kernel_ms_last, "last_step_split": e2e_split},
			"gpu_launches": launches + e2e_launches,
			"roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
						 "traffic": None, "algorithmic_bytes_per_launch": alg, "peak_source": peak_source,
						 "grid_barriers_per_launch": grid_barriers,
						 "note": "latency/barrier bound: see DESIGN.md (barrier floor) and profiles/"},
			"stage_ms_per_step": {n: stage_ms[i] / args.steps for i, n in enumerate(b2.STAGE_NAMES)},
			"clocks": clocks,
		}
		if cpu is not None:
			line["cpu_baseline"] = cpu
		print(json.dumps(line))
	if distributed:
		dist.destroy_process_group()
	return 0


def main() -> int:
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=100)
	ap.add_argument("--warmup", type=int, default=20)
	ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
	ap.add_argument("--scene", default="many_pyramids", choices=sorted(SCENE_STEPS))
	ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
	args = ap.parse_args()
	args.warmup = max(3, args.warmup)
	if args.impl == "reference":
		return run_reference_arm(args)
	return run_gpu_arm(args)


if __name__ == "__main__":
	sys.exit(main())
