// b2g_resident.cuh -- the two small kernels that keep the device-resident copies in step with the host (resident mode,
// b2g_types.cuh): what the reference leaves unchanged from one step to the next -- the manifolds it recycles
// (src/physics_world.c:508-560), the bodies' constants, the velocities the device computed itself -- does not cross PCIe
// again.
//
//   b2gApplyBodiesKernel   BEFORE the step's kernels, only when the pack pass found dirty bodies (a body the host touched:
//                          new in the awake set, moved by a swap-remove, velocity / force / mass set through the API):
//                          their records overwrite the resident state and constants.
//   b2gCommitKernel        AFTER the step's kernels, off the critical path (the download is already running): the static
//                          rows of the step's full records go into the table, at the contact's home, for the next step.
// Both are idempotent, so a step that is run again (b2GpuSolverRun twice, the rerun after binFail) sees the same inputs.
#pragma once

#include "b2g_contact.cuh"

namespace b2g
{

__global__ void __launch_bounds__( 256 ) b2gApplyBodiesKernel( const __grid_constant__ StepParams P )
{
	for ( int k = (int)( blockIdx.x * blockDim.x + threadIdx.x ); k < P.dirtyBodyCapacity; k += (int)( gridDim.x * blockDim.x ) )
	{
		const float4* record = P.dirtyBodies + (size_t)k * kDirtyBodyQuads;
		int body = __float_as_int( record[0].x );
		if ( body < 0 || body >= P.bodyCount )
		{
			continue; // the unused tail of a pack thread's chunk
		}
		P.residentStates[2 * (size_t)body + 0] = record[1];
		P.residentStates[2 * (size_t)body + 1] = record[2];
		P.residentBody[2 * (size_t)body + 0] = record[3];
		P.residentBody[2 * (size_t)body + 1] = record[4];
	}
}

__global__ void __launch_bounds__( 256 ) b2gCommitKernel( const __grid_constant__ StepParams P )
{
	// one thread per row of a slot: a warp moves 6.4 slots' worth of rows, coalesced on the source side.  Only the slots
	// of the colour ranges carry light records (the padding between the ranges is never written).
	for ( int c = 0; c <= P.colorCount; ++c )
	{
		const ColorRange range = c < P.colorCount ? P.colors[c] : P.overflow;
		const int rows = range.contactCount * kTableRows;
		for ( int t = (int)( blockIdx.x * blockDim.x + threadIdx.x ); t < rows; t += (int)( gridDim.x * blockDim.x ) )
		{
			int local = t / kTableRows, row = t - local * kTableRows;
			float4 L = P.light[range.contactStart + local];
			int key = __float_as_int( L.x ), ref = __float_as_int( L.w );
			if ( key >= 0 && ref < 0 )
			{
				float4 value = P.full[(size_t)( ~ref ) * WR_COUNT + row];
				if ( row == WR_HEAD )
				{
					value.w = 0.0f; // the rolling impulse is not a static row (it comes back with the output records)
				}
				P.table[(size_t)( key & kLightIdMask ) * kTableRows + row] = value;
			}
		}
	}
}

} // namespace b2g
