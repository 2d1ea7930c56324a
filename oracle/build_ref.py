#!/usr/bin/env python3
"""Recipe for oracle/_ref/: compile the UNTOUCHED reference (where it lies under /root/reference) with gcc into
shared libraries the tests and bench.py's cpu_baseline / --impl reference arm load.  Test infrastructure only.

    python oracle/build_ref.py            # libbox2d_ref.so, libbox2d_refcap.so, libbox2d_ref_avx2.so (+ liboracle.so)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tools import buildlib  # noqa: E402


def main() -> int:
	print(buildlib.build_oracle_lib(verbose=True))
	libs = buildlib.build_reference_libs(verbose=True)
	for name, path in libs.items():
		print(name, path)
	print(buildlib.build_reference_avx2(verbose=True))  # second CPU baseline for bench.py (8-wide path)
	return 0


if __name__ == "__main__":
	sys.exit(main())
