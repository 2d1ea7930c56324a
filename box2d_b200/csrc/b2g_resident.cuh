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

#include <cstddef>

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

// Joints in resident mode.  A prepared b2JointSim changes from step to step only in the run of fields b2PrepareJoint
// rewrites from the bodies' poses (b2lJointPreparedRun, include/b2gpu_layout.h: body indices, anchor frames, deltaCenter,
// the effective masses that depend on them) and in the accumulated impulses the device itself wrote.  So a joint that is
// otherwise unchanged travels as 96 bytes instead of 256: its home, where its previous output record is, and the run.
// This kernel rebuilds the complete records the stage code reads (P.rawJoints) and keeps the table up to date:
//     full record (ref < 0):   the record as uploaded
//     light record:            the table's record + the solver's own outputs of the previous step (the fields
//                              b2lJointMutableRuns lists) + the fresh prepared run, in that order -- the motor joint's
//                              linearMass is both an output and re-prepared, and the host's value is the prepared one
// Idempotent (the table ends up holding what was assembled), one thread per joint.
__global__ void __launch_bounds__( 256 ) b2gAssembleJointsKernel( const __grid_constant__ StepParams P )
{
	// 16 threads per joint, one per 16-byte quad of the record: coalesced on every side, no per-thread record
	constexpr int quads = kJointStride / 16;
	const int total = P.jointCount * quads;
	for ( int t = (int)( blockIdx.x * blockDim.x + threadIdx.x ); t < total; t += (int)( gridDim.x * blockDim.x ) )
	{
		const int j = t / quads, q = t - j * quads;
		const float4* light = P.lightJoints + (size_t)j * kLightJointQuads;
		const float4 head = light[0];
		const int home = __float_as_int( head.x ), ref = __float_as_int( head.y );
		float4* out = P.jointAssembled + (size_t)j * quads;
		float4* keep = P.jointTable + (size_t)home * quads;
		if ( ref < 0 )
		{
			float4 value = P.fullJoints[(size_t)( ~ref ) * quads + q];
			out[q] = value;
			keep[q] = value;
			continue;
		}
		float4 value = keep[q];
		const int type = __float_as_int( keep[0].w ); // b2JointSim::type, the fourth word of the record
		static_assert( offsetof( b2lJointSim, type ) == 12, "type is expected in the first quad" );
		float words[4] = { value.x, value.y, value.z, value.w };
		int offsets[2], floats[2];
		const int runs = b2lJointMutableRuns( type, offsets, floats );
		int runOffset = 0;
		const int runBytes = b2lJointPreparedRun( type, &runOffset );
		const float* previous = P.prevOutJoints + (size_t)ref * B2L_JOINT_OUT_FLOATS;
		const float* run = reinterpret_cast<const float*>( light + 1 );
#pragma unroll
		for ( int i = 0; i < 4; ++i )
		{
			const int word = 4 * q + i; // index of the float in the record
			// the solver's own outputs of the previous step first ...
			int before = 0;
			for ( int r = 0; r < runs; ++r )
			{
				int first = offsets[r] / 4;
				if ( word >= first && word < first + floats[r] )
				{
					words[i] = previous[before + word - first];
				}
				before += floats[r];
			}
			// ... then what b2PrepareJoint wrote this step (it wins where the two overlap: the motor joint's linearMass)
			int first = runOffset / 4;
			if ( word >= first && word < first + runBytes / 4 )
			{
				words[i] = run[word - first];
			}
		}
		value = make_float4( words[0], words[1], words[2], words[3] );
		out[q] = value;
		keep[q] = value;
	}
}

} // namespace b2g
