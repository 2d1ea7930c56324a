import ctypes, numpy as np, sys
sys.path.insert(0,'/root/repo')
import box2d_b200 as b2
host=b2.host_lib(); host.b2GpuSeam_InstallPinnedAllocator()
for scene in ("many_pyramids","rain","large_pyramid","joint_grid","tumbler"):
    with b2.World(host, scene, 8) as w:
        w.step(30 if scene!="rain" else 200)
        d=host.b2GpuSeam_GetLastDesc(w.world_index()).contents
        r=host.b2GpuSeam_GetLastResult(w.world_index()).contents
        n=d.awakeBodyCount
        lab=np.ctypeslib.as_array(ctypes.cast(d.bodyIsland, ctypes.POINTER(ctypes.c_int)), shape=(n,)).copy()
        cnt=np.bincount(lab[lab>=0], minlength=d.islandCount)
        print(scene,"bodies",n,"islands",d.islandCount,"neg",int((lab<0).sum()),"max island",cnt.max(),"launches",r.kernelLaunches,"barriers",r.gridBarriers,"kernel ms",r.kernelMs, "overflow", d.overflow.contactCount, d.overflow.jointCount)
