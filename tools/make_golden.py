#!/usr/bin/env python3
"""Generate tests/golden/*.b2cap.gz: kernel-level golden vectors captured from the UNTOUCHED reference solver.

Runs here (needs /root/reference to build oracle/_ref/libbox2d_refcap.so).  Each capture holds the exact inputs the
C-ABI receives for one solver step and the outputs the reference's CPU solver produced (oracle/harness/b2h_capture.c).
Also writes tests/golden/hashes.json: b2World_GetStateHash of the reference after N steps per scene.

    python tools/make_golden.py
"""
import ctypes
import gzip
import json
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tools import buildlib  # noqa: E402
import box2d_b200 as b2  # noqa: E402

# scene, steps to run before the captured step (small scenes only: these are committed fixtures)
CAPTURES = [
	("small_pyramid", 0), ("small_pyramid", 30),
	("pyramid_soft", 5), ("pyramid_cold", 3),
	("falling_hinges", 20), ("falling_hinges", 120),
	("joint_zoo", 0), ("joint_zoo", 45), ("joint_zoo_cold", 10),
	("contact_zoo", 2), ("contact_zoo", 40), ("contact_zoo", 130),
	("overflow", 1), ("overflow", 25),
]

# scene -> steps for the state-hash goldens (the first five are SURVEY.md section 8c's measured values)
HASHES = [
	("large_pyramid", 100), ("many_pyramids", 40), ("joint_grid", 100), ("rain", 300), ("tumbler", 200),
	("small_pyramid", 201), ("joint_zoo", 120), ("contact_zoo", 200), ("overflow", 120), ("pyramid_soft", 60),
	("pyramid_cold", 60), ("joint_zoo_cold", 60), ("spinner", 60), ("smash", 60), ("compounds", 60), ("washer", 40),
]


def main() -> int:
	libs = buildlib.build_reference_libs()
	cap = b2._bind_harness(ctypes.CDLL(str(libs["refcap"])))
	cap.b2h_capture_arm.argtypes = [ctypes.c_char_p]
	golden = ROOT / "tests" / "golden"
	golden.mkdir(parents=True, exist_ok=True)

	for scene, before in CAPTURES:
		with b2.World(cap, scene, workers=2) as w:
			w.step(before)
			with tempfile.TemporaryDirectory() as tmp:
				raw = os.path.join(tmp, "cap.bin")
				cap.b2h_capture_arm(raw.encode())
				w.step(1)
				data = open(raw, "rb").read()
			out = golden / f"{scene}_{before:03d}.b2cap.gz"
			with gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as f:
				f.write(data)
			c = b2.Capture(out)
			print(f"{out.name}: bodies {c.body_count} contacts {c.contact_count} joints {c.joint_count} "
				  f"colours {c.desc.activeColorCount} overflow {c.color_counts[-1]} ({out.stat().st_size} B)")

	ref = b2._bind_harness(ctypes.CDLL(str(libs["ref"])))
	hashes = {}
	for scene, steps in HASHES:
		with b2.World(ref, scene, workers=4) as w:
			w.step(steps)
			hashes[scene] = {"steps": steps, "hash": f"{w.hash():016x}"}
		print(scene, hashes[scene])
	with b2.World(ref, "falling_hinges", workers=1) as w:
		w.step(300)
		done, sleep_step, h = w.hinges_result()
		hashes["falling_hinges"] = {"sleepStep": sleep_step, "transformHash": f"{h:08x}"}
	(golden / "hashes.json").write_text(json.dumps(hashes, indent=1) + "\n")
	return 0


if __name__ == "__main__":
	sys.exit(main())
