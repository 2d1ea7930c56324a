"""Probe: N live worlds in a group, time per round of steps (GPU box)."""
import ctypes, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import box2d_b200 as b2
n = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
gpu = b2.host_lib(); gpu.b2GpuSeam_InstallPinnedAllocator()
t = time.time(); ws = [b2.World(gpu, "small_pyramid", 1, variant=1 + i) for i in range(n)]; print("create", time.time() - t)
with b2.WorldGroup(gpu, ws) as g:
    g.step(30)
    t = time.time(); g.step(steps); dt = time.time() - t
    print(f"{n} live worlds: {dt / steps * 1e3:.2f} ms per batch step (wall, whole b2World_Step of every world)")
    r = gpu.b2GpuSeam_GetLastResult(ws[0].world_index()).contents
    print("last batch: kernel %.3f ms abi %.3f ms pack %.3f wait %.3f unpack %.3f" % (r.kernelMs, r.totalMs, r.uploadMs, r.waitMs, r.scatterMs))
t = time.time(); b2.step_many(gpu, ws[:112], 5); print("ungrouped sanity", time.time() - t)
