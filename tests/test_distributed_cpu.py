"""The N > 1 path of bench.py on CPU: world_size-2 gloo processes shard the independent worlds with no exchange and
reduce only the result (max of the timings over ranks, sum of the work), exactly what the NCCL run does on GPUs."""
import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port() -> int:
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _worker(rank: int, world_size: int, port: int, total: int, out_dir: str):
	import torch.distributed as dist

	sys.path.insert(0, str(ROOT))
	import bench

	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world_size)
	begin, end = bench.shard_range(total, rank, world_size)
	# pretend rank r needs (r + 1) seconds for its share of the worlds
	(kernel_s, e2e_s), work = bench.reduce_over_ranks([1.0 + rank, 10.0 - rank], float(end - begin))
	Path(out_dir, f"rank{rank}.txt").write_text(f"{begin} {end} {kernel_s} {e2e_s} {work}")
	dist.barrier()
	dist.destroy_process_group()


def test_two_ranks_shard_and_reduce(tmp_path):
	import torch.multiprocessing as mp

	world_size, total = 2, 8193
	mp.spawn(_worker, args=(world_size, _free_port(), total, str(tmp_path)), nprocs=world_size, join=True)
	rows = [Path(tmp_path, f"rank{r}.txt").read_text().split() for r in range(world_size)]
	ranges = [(int(r[0]), int(r[1])) for r in rows]
	assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == total  # disjoint cover, no exchange
	assert abs((ranges[0][1] - ranges[0][0]) - (ranges[1][1] - ranges[1][0])) <= 1
	for r in rows:
		assert float(r[2]) == 2.0  # max over ranks
		assert float(r[3]) == 10.0
		assert float(r[4]) == float(total)  # sum of the work


def test_shard_range_properties():
	import bench

	for total in (0, 1, 7, 8192, 8193):
		for n in (1, 2, 4, 8):
			parts = [bench.shard_range(total, r, n) for r in range(n)]
			assert parts[0][0] == 0 and parts[-1][1] == total
			assert all(parts[i][1] == parts[i + 1][0] for i in range(n - 1))
			sizes = [e - b for b, e in parts]
			assert max(sizes) - min(sizes) <= 1


def test_single_process_reduce_is_identity():
	import bench

	(a, b), w = bench.reduce_over_ranks([0.25, 0.5], 123.0)
	assert (a, b, w) == (0.25, 0.5, 123.0)
