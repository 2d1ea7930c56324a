#!/usr/bin/env python3
"""Build the product: libb2gpusolver.so (CUDA, sm_100a) and libbox2d_b200.so (reference host + seam)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tools import buildlib  # noqa: E402


def main() -> int:
	print(buildlib.build_cuda_lib(verbose="-v" in sys.argv))
	if buildlib.reference_available():
		print(buildlib.build_host_lib(verbose=True))
	else:
		print("reference sources absent: keeping the prebuilt host library")
	return 0


if __name__ == "__main__":
	sys.exit(main())
