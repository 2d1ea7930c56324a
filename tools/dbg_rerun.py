import os, sys, numpy as np
os.environ["B2GPU_TEST_TIGHT_BINS"]="1"
sys.path.insert(0,'.')
import box2d_b200 as b2
files=sorted((b2.ROOT/"tests"/"golden").glob("*.b2cap.gz"))
with b2.GpuSolver() as solver:
    for path in files:
        cap=b2.Capture(path)
        desc,result,bufs=cap.make_call(islands=True)
        solver.step(desc,result)
        ok=np.array_equal(bufs["states"].view(np.uint32), cap.states_out.view(np.uint32))
        nbad=int((bufs["states"].view(np.uint32)!=cap.states_out.view(np.uint32)).any(axis=-1).sum()) if not ok else 0
        print(path.name, "ok" if ok else "BAD %d/%d"%(nbad,len(cap.states_out)), "plan",solver.island_plan(),"gb",result.gridBarriers,"launches",result.kernelLaunches)
