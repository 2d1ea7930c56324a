/*
 * b2o_solver.h -- CPU oracle of the Soft Step solver (TEST INFRASTRUCTURE).
 *
 * A plain-C, single-threaded restatement of everything b2SolverTask runs (reference src/solver.c:1055-1197) over the
 * SAME descriptor the product's C-ABI takes (include/b2_gpu_solver.h), so a test can hand identical inputs to the CUDA
 * path and to this file and compare bits.  It is pinned against the reference itself: tests/test_oracle_cpu.py checks
 * it bit-for-bit against the golden captures taken from the untouched reference (tests/golden/*.b2cap.gz, produced by
 * tools/make_golden.py through oracle/_ref/libbox2d_refcap.so) -- parity pinned, see DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import, link or execute anything under
 * oracle/.  The product (box2d_b200/) never does and fails loudly without its CUDA extension.
 */
#ifndef B2O_SOLVER_H
#define B2O_SOLVER_H

#include "b2_gpu_solver.h"
#include "b2gpu_layout.h"
#include "b2o_math.h"

/* b2BodyState, reference src/body.h:153-168 */
typedef struct
{
	o_vec2 v;
	float w;
	uint32_t flags;
	o_vec2 dp;
	o_rot dq;
} o_state;

typedef struct
{
	o_state* states;
	float h, inv_h, inv_dt;
	float lengthUnitsPerMeter;
} o_ctx;

void b2o_warm_start_joint( b2lJointSim* joint, const o_ctx* ctx );
void b2o_solve_joint( b2lJointSim* joint, const o_ctx* ctx, bool useBias );
void b2o_joint_reaction( const b2lJointSim* sim, float invTimeStep, float* force, float* torque );

/* Same contract as b2GpuSolverStep: solves in place on the descriptor's host arrays. Returns 0. */
__attribute__( ( visibility( "default" ) ) ) int b2OracleSolverStep( const b2GpuStepDesc* desc, b2GpuStepResult* result );

#endif
