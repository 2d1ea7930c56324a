"""Mean b2Profile of a scene over a number of steps, GPU host library vs the reference (ms per step).
   python tools/profile_steps.py <scene> [steps] [workers]"""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import box2d_b200 as b2

scene = sys.argv[1] if len(sys.argv) > 1 else "many_pyramids"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
workers = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ref = b2._bind_harness(ctypes.CDLL(str(b2.ROOT / "oracle" / "_ref" / "libbox2d_ref.so")))
gpu = b2.host_lib()
gpu.b2GpuSeam_InstallPinnedAllocator()
names = ("step", "pairs", "collide", "solve", "solverSetup", "constraints", "transforms", "refit", "bullets", "sleepIslands", "jointEvents", "hitEvents", "sensors")
for label, lib in (("reference", ref), ("gpu", gpu)):
	with b2.World(lib, scene, workers) as w:
		w.step(30)
		acc = {n: 0.0 for n in names}
		for _ in range(steps):
			w.step()
			p = w.profile()
			for n in names:
				acc[n] += p[n] / steps
		print(label, " ".join(f"{n}={acc[n]:.3f}" for n in names))
