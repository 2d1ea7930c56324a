"""Build recipes shared by oracle/build_ref.py (reference + oracle) and box2d_b200/build.py (product).

Nothing from /root/reference is copied into the repository: reference translation units are compiled from
where they lie, object files and generated (patched) sources go to git-ignored directories
(oracle/_ref/, box2d_b200/host/_gen/), and only shared libraries come out.

We do not run the reference's own build system; the flags below restate what its CMake sets for a Release
build with gcc (reference CMakeLists.txt:50-63 -ffp-contract=off, src/CMakeLists.txt:104-111 C17,
CMake's Release default -O3 -DNDEBUG; SSE2 width 4 is the default x86-64 path, src/core.h:50-75).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path(os.environ.get("B2_REFERENCE", "/root/reference"))
REF_OUT = ROOT / "oracle" / "_ref"
OBJ_DIR = REF_OUT / "obj"
GEN_DIR = ROOT / "box2d_b200" / "host" / "_gen"
PKG_DIR = ROOT / "box2d_b200"

CC = os.environ.get("CC", "gcc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

REF_CFLAGS = ["-O3", "-DNDEBUG", "-std=gnu17", "-ffp-contract=off", "-fPIC", "-w"]
# The product library can hold a whole batch of live worlds (b2GpuSeam_CreateGroup): B2_MAX_WORLDS is a documented knob of the
# reference (include/box2d/constants.h:38-40, "must stay < 65535"); it only sizes the static world array of
# src/physics_world.c and the seam's slot table.  The reference builds under oracle/_ref keep the default (128).
PRODUCT_DEFINES = ["-DB2_MAX_WORLDS=8192"]
OWN_CFLAGS = ["-O2", "-DNDEBUG", "-std=gnu17", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra"]

NVCC_FLAGS = [
	"-gencode", "arch=compute_100a,code=sm_100a",
	"-O3", "-lineinfo", "-std=c++17",
	# bit-exactness with the reference's -ffp-contract=off host build (SURVEY.md section 8a parity notes)
	"-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
	"-Xcompiler", "-fPIC",
]


def reference_available() -> bool:
	return (REFERENCE / "src" / "solver.c").is_file()


def _run(cmd: list[str], quiet: bool = True) -> None:
	proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
	if proc.returncode != 0:
		sys.stderr.write(" ".join(str(c) for c in cmd) + "\n" + proc.stdout + "\n")
		raise RuntimeError(f"command failed: {cmd[0]} ... {cmd[-1]}")
	if not quiet and proc.stdout.strip():
		print(proc.stdout)


def _stale(target: Path, sources: list[Path]) -> bool:
	if not target.exists():
		return True
	t = target.stat().st_mtime
	return any(s.exists() and s.stat().st_mtime > t for s in sources)


def ref_includes() -> list[str]:
	return [f"-I{REFERENCE / 'include'}", f"-I{REFERENCE / 'src'}", f"-I{REFERENCE / 'shared'}"]


def own_includes() -> list[str]:
	return [f"-I{ROOT / 'include'}", f"-I{PKG_DIR / 'host'}"]


def compile_reference_objects() -> list[Path]:
	"""Every reference src/*.c except solver.c, plus shared/*.c, compiled in place -> oracle/_ref/obj/*.o"""
	OBJ_DIR.mkdir(parents=True, exist_ok=True)
	jobs = []
	for src in sorted((REFERENCE / "src").glob("*.c")):
		if src.name == "solver.c":
			continue
		jobs.append((src, OBJ_DIR / f"src_{src.stem}.o"))
	for src in sorted((REFERENCE / "shared").glob("*.c")):
		jobs.append((src, OBJ_DIR / f"shared_{src.stem}.o"))

	def one(job):
		src, obj = job
		if _stale(obj, [src]):
			_run([CC, *REF_CFLAGS, *ref_includes(), "-c", str(src), "-o", str(obj)])
		return obj

	with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
		return list(pool.map(one, jobs))


def compile_solver_variant(mode: str) -> Path:
	"""mode: 'pure' (reference solver.c as is), 'hook' (capture hooks), 'gpu' (seam call)."""
	OBJ_DIR.mkdir(parents=True, exist_ok=True)
	src = REFERENCE / "src" / "solver.c"
	obj = OBJ_DIR / f"solver_{mode}.o"
	if mode == "pure":
		if _stale(obj, [src]):
			_run([CC, *REF_CFLAGS, *ref_includes(), "-c", str(src), "-o", str(obj)])
		return obj
	gen_dir = GEN_DIR if mode == "gpu" else REF_OUT / "gen"
	gen_dir.mkdir(parents=True, exist_ok=True)
	gen = gen_dir / f"solver_{mode}.c"
	patcher = ROOT / "tools" / "patch_solver.py"
	if _stale(gen, [src, patcher]):
		_run([sys.executable, str(patcher), "--mode", mode, "--src", str(src), "--out", str(gen)])
	if _stale(obj, [gen]):
		_run([CC, *REF_CFLAGS, *ref_includes(), "-c", str(gen), "-o", str(obj)])
	return obj


def compile_collide_gpu() -> Path:
	"""The product's physics_world.c: the reference's, with the two recycle hooks of tools/patch_collide.py."""
	OBJ_DIR.mkdir(parents=True, exist_ok=True)
	GEN_DIR.mkdir(parents=True, exist_ok=True)
	src = REFERENCE / "src" / "physics_world.c"
	gen = GEN_DIR / "physics_world_gpu.c"
	obj = OBJ_DIR / "physics_world_gpu.o"
	patcher = ROOT / "tools" / "patch_collide.py"
	if _stale(gen, [src, patcher]):
		_run([sys.executable, str(patcher), "--src", str(src), "--out", str(gen)])
	if _stale(obj, [gen, Path(__file__)]):
		_run([CC, *REF_CFLAGS, *PRODUCT_DEFINES, *ref_includes(), "-c", str(gen), "-o", str(obj)])
	return obj


# Deferred impulses (box2d_b200/host/b2_gpu_seam.c): the reference's readers of manifold impulses outside the narrow phase and the
# solver.  In the PRODUCT's copies of the reference's object files their definitions are renamed to b2Ref_* (objcopy; the sources
# are not touched, the reference builds under oracle/_ref keep the originals) and the seam defines the public names: a flush of
# the pending impulses, then the reference's function.
INTERPOSED = {
	"src_body.o": ["b2Body_GetContactData"],
	"src_shape.o": ["b2Shape_GetContactData"],
	"src_contact.o": ["b2Contact_GetData"],
	"physics_world_gpu.o": ["b2World_Draw"],
	"src_world_snapshot.o": ["b2World_GetStateHash", "b2World_Snapshot", "b2World_Restore", "b2SerializeWorld", "b2HashWorldStateDeep"],
	"src_solver_set.o": ["b2TrySleepIsland", "b2TransferJoint"],
	"src_joint.o": ["b2GetJointSimCheckType", "b2Joint_GetConstraintForce", "b2Joint_GetConstraintTorque"],
	# ... and the two functions that change a colour's contact array between the narrow phase and the solver (recycledInPlace)
	"src_constraint_graph.o": ["b2AddContactToGraph", "b2RemoveContactFromGraph", "b2RemoveJointFromGraph"],
}
OBJCOPY = os.environ.get("OBJCOPY", "objcopy")


def interpose(obj: Path) -> Path:
	"""The product's copy of a reference object file, with the INTERPOSED definitions renamed to b2Ref_*."""
	names = INTERPOSED.get(obj.name)
	if not names:
		return obj
	out = obj.with_name(obj.stem + "_interposed.o")
	if _stale(out, [obj, Path(__file__)]):
		cmd = [OBJCOPY]
		for name in names:
			cmd += ["--redefine-sym", f"{name}=b2Ref_{name[2:]}"]
		_run(cmd + [str(obj), str(out)])
		# every name must have been defined there, or the seam's wrapper would call itself
		syms = subprocess.run(["nm", "--defined-only", str(out)], stdout=subprocess.PIPE, text=True, check=True).stdout
		for name in names:
			if f" T b2Ref_{name[2:]}\n" not in syms:
				out.unlink()
				raise RuntimeError(f"interpose: {obj.name} does not define {name}")
	return out


def compile_own_c(src: Path, tag: str = "", defines: list[str] | None = None) -> Path:
	OBJ_DIR.mkdir(parents=True, exist_ok=True)
	obj = OBJ_DIR / f"own_{src.stem}{tag}.o"
	headers = list((ROOT / "include").glob("*.h")) + list((PKG_DIR / "host").glob("*.h"))
	if _stale(obj, [src, *headers, Path(__file__)]):
		_run([CC, *OWN_CFLAGS, *(defines or []), *own_includes(), *ref_includes(), "-c", str(src), "-o", str(obj)])
	return obj


def link_shared(target: Path, objects: list[Path], extra: list[str] | None = None) -> None:
	target.parent.mkdir(parents=True, exist_ok=True)
	if _stale(target, objects):
		cmd = [CC, "-shared", "-o", str(target), *[str(o) for o in objects], "-Wl,-Bsymbolic", "-lm", "-lpthread"]
		_run(cmd + (extra or []))


def build_reference_libs(verbose: bool = False) -> dict[str, Path]:
	"""oracle/_ref/libbox2d_ref.so (untouched reference + scene harness) and libbox2d_refcap.so (capture)."""
	if not reference_available():
		raise RuntimeError(f"reference sources not found at {REFERENCE}")
	objs = compile_reference_objects()
	harness = compile_own_c(PKG_DIR / "host" / "b2h_harness.c")
	pure = compile_solver_variant("pure")
	hook = compile_solver_variant("hook")
	capture = compile_own_c(ROOT / "oracle" / "harness" / "b2h_capture.c")
	seam_desc = compile_own_c(PKG_DIR / "host" / "b2_gpu_seam_desc.c")

	ref = REF_OUT / "libbox2d_ref.so"
	refcap = REF_OUT / "libbox2d_refcap.so"
	link_shared(ref, [*objs, pure, harness])
	link_shared(refcap, [*objs, hook, harness, capture, seam_desc])
	if verbose:
		print(f"built {ref} and {refcap}")
	return {"ref": ref, "refcap": refcap}


def build_cuda_lib(verbose: bool = False) -> Path:
	"""box2d_b200/libb2gpusolver.so: the CUDA kernels + the C-ABI of include/b2_gpu_solver.h (sm_100a only)."""
	csrc = PKG_DIR / "csrc"
	sources = sorted(csrc.glob("*.cu"))
	headers = sorted(csrc.glob("*.cuh")) + sorted(csrc.glob("*.h")) + sorted((ROOT / "include").glob("*.h"))
	target = PKG_DIR / "libb2gpusolver.so"
	if _stale(target, [*sources, *headers]):
		cmd = [NVCC, *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{csrc}", "-shared", "-o", str(target),
			   *[str(s) for s in sources]]
		if verbose:
			cmd.insert(1, "-Xptxas")
			cmd.insert(2, "-v")
		_run(cmd, quiet=not verbose)
	return target


def build_host_lib(verbose: bool = False) -> Path:
	"""box2d_b200/libbox2d_b200.so: the reference's host code with the solve region replaced by the seam."""
	if not reference_available():
		raise RuntimeError(f"reference sources not found at {REFERENCE}")
	cuda_lib = build_cuda_lib(verbose)
	objs = [interpose(o) for o in compile_reference_objects() if o.name != "src_physics_world.o"] + [interpose(compile_collide_gpu())]
	gpu = compile_solver_variant("gpu")
	harness = compile_own_c(PKG_DIR / "host" / "b2h_harness.c")
	seam = compile_own_c(PKG_DIR / "host" / "b2_gpu_seam.c", "_product", PRODUCT_DEFINES)
	seam_desc = compile_own_c(PKG_DIR / "host" / "b2_gpu_seam_desc.c", "_product", PRODUCT_DEFINES)
	target = PKG_DIR / "libbox2d_b200.so"
	link_shared(target, [*objs, gpu, harness, seam, seam_desc, cuda_lib],
				[f"-L{PKG_DIR}", "-lb2gpusolver", "-Wl,-rpath,$ORIGIN"])
	if verbose:
		print(f"built {target}")
	return target


def build_reference_avx2(verbose: bool = False) -> Path:
	"""oracle/_ref/libbox2d_ref_avx2.so: the untouched reference with its optional 8-wide path (BOX2D_AVX2, -mavx2: reference
	src/CMakeLists.txt:187-190).  Only a second CPU baseline for bench.py -- not bit-compatible with the default build (the
	reference documents the wrappers' differences, src/contact_solver.c:641-645) and not used as a checker."""
	if not reference_available():
		raise RuntimeError(f"reference sources not found at {REFERENCE}")
	out = REF_OUT / "obj_avx2"
	out.mkdir(parents=True, exist_ok=True)
	flags = [*REF_CFLAGS, "-DBOX2D_AVX2", "-mavx2"]
	jobs = [(src, out / f"src_{src.stem}.o") for src in sorted((REFERENCE / "src").glob("*.c"))]
	jobs += [(src, out / f"shared_{src.stem}.o") for src in sorted((REFERENCE / "shared").glob("*.c"))]
	harness = PKG_DIR / "host" / "b2h_harness.c"
	jobs.append((harness, out / "own_b2h_harness.o"))

	def one(job):
		src, obj = job
		if _stale(obj, [src]):
			_run([CC, *flags, *own_includes(), *ref_includes(), "-c", str(src), "-o", str(obj)])
		return obj

	with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
		objs = list(pool.map(one, jobs))
	target = REF_OUT / "libbox2d_ref_avx2.so"
	link_shared(target, objs)
	if verbose:
		print(f"built {target}")
	return target


def build_oracle_lib(verbose: bool = False) -> Path:
	"""oracle/liboracle.so: the plain-C restatement of the solver (checker only, needs nothing from the reference)."""
	odir = ROOT / "oracle"
	sources = [odir / "b2o_solver.c", odir / "b2o_joints.c"]
	headers = [odir / "b2o_solver.h", odir / "b2o_math.h", *sorted((ROOT / "include").glob("*.h"))]
	target = odir / "liboracle.so"
	if _stale(target, [*sources, *headers]):
		_run([CC, "-O2", "-std=gnu17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra", "-Wno-comment",
			  f"-I{ROOT / 'include'}", f"-I{odir}", *[str(s) for s in sources], "-lm", "-o", str(target)])
	if verbose:
		print(f"built {target}")
	return target


def clean() -> None:
	for d in (REF_OUT, GEN_DIR):
		shutil.rmtree(d, ignore_errors=True)
	for so in PKG_DIR.glob("*.so"):
		so.unlink()
