#!/usr/bin/env python3
"""How well do owner lists work for a scene?  Steps a scene with the untouched reference (CPU only), captures one solver
step and reports, for clusters of 8 and 16 blocks with equal body runs: the largest block's share of the contacts
relative to the average (capacity head room needed) and the fraction of contacts whose bodies both live in the block
that owns the first one (gathers / scatters that stay in the block's own shared memory).
   python tools/owner_balance.py [scene steps] ...        default: large_pyramid 60, tumbler 130"""
import ctypes
import gzip
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import box2d_b200 as b2  # noqa: E402
import buildlib  # noqa: E402


def main() -> int:
	args = sys.argv[1:]
	scenes = [(args[i], int(args[i + 1])) for i in range(0, len(args) - 1, 2)] or [("large_pyramid", 60), ("tumbler", 130)]
	libs = buildlib.build_reference_libs()
	cap = b2._bind_harness(ctypes.CDLL(str(libs["refcap"])))
	cap.b2h_capture_arm.argtypes = [ctypes.c_char_p]
	for scene, steps in scenes:
		with b2.World(cap, scene, workers=2) as w:
			w.step(steps)
			raw = os.path.join(tempfile.mkdtemp(), "capture.bin")
			cap.b2h_capture_arm(raw.encode())
			w.step(1)
		packed = raw + ".gz"
		with gzip.GzipFile(packed, "wb") as f:
			f.write(open(raw, "rb").read())
		c = b2.Capture(packed)
		n = c.body_count
		for blocks in (8, 16):
			run = ((n + blocks - 1) // blocks + 3) & ~3
			share = np.zeros(blocks, int)
			local = total = 0
			for colour in c.contacts_in[:-1]:  # the overflow colour goes to the first block as a whole
				if colour.size == 0:
					continue
				sims = colour.reshape(-1, b2.CONTACT_SIZE)
				a = sims[:, 36:40].copy().view(np.int32).ravel()
				b = sims[:, 40:44].copy().view(np.int32).ravel()
				owner = np.where(a >= 0, a, b) // run
				share += np.bincount(owner, minlength=blocks)
				in_a = (a < 0) | (a // run == owner)
				in_b = (b < 0) | (b // run == owner)
				local += int((in_a & in_b).sum())
				total += len(a)
			print(f"{scene}: {n} bodies, {total} coloured contacts, {blocks} blocks x {run} bodies: largest share "
				  f"{share.max()} = {share.max() / (total / blocks):.2f} x average, both bodies local for {local / total:.0%}")
	return 0


if __name__ == "__main__":
	sys.exit(main())
