// b2g_batch.cu -- batch of independent worlds (placeholder until the world-per-block kernel lands).
#include "b2_gpu_solver.h"

extern "C" int b2GpuSolverUploadBatch( b2GpuSolver*, const b2GpuStepDesc*, int )
{
	return 1;
}
extern "C" int b2GpuSolverRunBatch( b2GpuSolver*, b2GpuStepResult* )
{
	return 1;
}
extern "C" int b2GpuSolverDownloadBatch( b2GpuSolver*, const b2GpuStepDesc*, int, b2GpuStepResult* )
{
	return 1;
}
extern "C" int b2GpuSolverStepBatch( b2GpuSolver*, const b2GpuStepDesc*, int, b2GpuStepResult* )
{
	return 1;
}
