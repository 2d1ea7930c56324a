"""Step a benchmark scene through the product library and print, every 10 steps, what the seam handed to the device and
how the device solved it (island plan, launches, grid barriers, kernel and ABI time, stage split).
   python tools/scene_trace.py <scene> [blocks of 10 steps]      (needs a GPU; scenes: box2d_b200/host/b2h_harness.c)"""
import ctypes, numpy as np, sys
sys.path.insert(0,'.')
import box2d_b200 as b2
host=b2.host_lib(); host.b2GpuSeam_InstallPinnedAllocator()
scene=sys.argv[1] if len(sys.argv)>1 else "tumbler"
with b2.World(host, scene, 8) as w:
    for i in range(int(sys.argv[2]) if len(sys.argv)>2 else 16):
        w.step(10)
        d=host.b2GpuSeam_GetLastDesc(w.world_index()).contents
        r=host.b2GpuSeam_GetLastResult(w.world_index()).contents
        n=d.awakeBodyCount
        lab=np.ctypeslib.as_array(ctypes.cast(d.bodyIsland, ctypes.POINTER(ctypes.c_int)), shape=(n,)).copy()
        cnt=np.bincount(lab[lab>=0], minlength=max(1,d.islandCount))
        contacts=sum(d.colors[c].contactCount for c in range(d.activeColorCount))
        print("step",(i+1)*10,"bodies",n,"islands",d.islandCount,"max island",cnt.max(),"colors",d.activeColorCount,"contacts",contacts,"overflow",d.overflow.contactCount,d.overflow.jointCount,"| launches",r.kernelLaunches,"barriers",r.gridBarriers,"kernel ms %.3f"%r.kernelMs,"total ms %.3f"%r.totalMs, "stages", ["%.3f"%r.stageMs[k] for k in range(8)])
