#!/usr/bin/env python3
"""Turn the ncu artefacts of tools/profile_round.sh (gpurun_out/<round>_*) into the tracked summaries under profiles/:
   <round>_launches_default.csv      the launch list as ncu wrote it
   <round>_launch_summary.txt        per kernel: launches, mean/min duration, share of the step
   <round>_<kernel>_raw.csv          the --page raw metrics that the roofline numbers quote
   <round>_<kernel>_summary.txt      duration, DRAM bytes, stall reasons, occupancy
   <round>_bench_*.json              the bench lines of the same box"""
import csv, io, json, shutil, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit", "sm__cycles_elapsed.max", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg")


def launch_summary(rnd: str) -> None:
	src = OUT / f"{rnd}_launches_default.csv"
	if not src.exists():
		return
	shutil.copy(src, PROF / src.name)
	rows = list(csv.reader(open(src)))
	header, agg = None, {}
	for r in rows:
		if len(r) > 10 and r[0] == "ID":
			header = r
			continue
		if header and len(r) == len(header):
			d = dict(zip(header, r))
			if d["Metric Name"] != "gpu__time_duration.sum":
				continue
			v = float(d["Metric Value"].replace(",", ""))
			v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
			agg.setdefault(d["Kernel Name"].split("(")[0], []).append(v)
	ours = {k: v for k, v in agg.items() if k.startswith("b2g")}
	steps = max(len(v) for v in ours.values()) if ours else 1
	total = sum(sum(v) for v in ours.values()) / steps
	with open(PROF / f"{rnd}_launch_summary.txt", "w") as f:
		f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 3 --warmup 3 --no-batch --no-cpu-baseline (many_pyramids)\n")
		f.write("per-launch times under ncu are serialised and cold-cache: the SHARE of the step is what must agree with bench.py\n\n")
		f.write(f"{'kernel':42s} {'launches':>8s} {'mean us':>9s} {'min us':>9s} {'share of step':>14s}\n")
		for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
			share = f"{100.0 * (sum(v) / steps) / total:13.1f}%" if k in ours else "   (not ours)"
			f.write(f"{k[:42]:42s} {len(v):8d} {sum(v) / len(v):9.2f} {min(v):9.2f} {share}\n")
		f.write(f"\nkernels of one solver step (sum of our kernels / steps): {total:.2f} us\n")


def kernel_summary(rnd: str, name: str, note: str) -> None:
	rep = OUT / f"{rnd}_{name}.ncu-rep"
	page = OUT / f"{rnd}_{name}.rawpage.csv"  # written on the GPU box by tools/profile_round.sh (the reports are too big to bring back)
	if page.exists() and page.stat().st_size > 0:
		raw = page.read_text()
	elif rep.exists():
		raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
	else:
		return
	rows = list(csv.reader(io.StringIO(raw)))
	if len(rows) < 3:
		return
	header, units, values = rows[0], rows[1], rows[-1]
	table = [(h, u, v) for h, u, v in zip(header, units, values)]
	with open(PROF / f"{rnd}_{name}_raw.csv", "w", newline="") as f:
		w = csv.writer(f)
		w.writerow(["metric", "unit", "value"])
		for h, u, v in table:
			if h in ("Kernel Name", "Block Size", "Grid Size") or h.startswith(KEEP):
				w.writerow([h, u, v])
	d = {h: v for h, _, v in table}
	u = {h: x for h, x, _ in table}

	def num(key: str) -> float:
		try:
			return float(d[key].replace(",", ""))
		except (KeyError, ValueError):
			return float("nan")

	def in_bytes(key: str) -> float:
		return num(key) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u.get(key, "byte"), 1.0)

	stalls = sorted(((num(k), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in d
					 if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and num(k) == num(k)), reverse=True)
	total = sum(v for v, _ in stalls) or 1.0
	dur = num("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u.get("gpu__time_duration.sum", "us"), 1.0)
	rd, wr = in_bytes("dram__bytes_read.sum"), in_bytes("dram__bytes_write.sum")
	with open(PROF / f"{rnd}_{name}_summary.txt", "w") as f:
		f.write(f"{d.get('Kernel Name', name)}   [{note}]\n")
		f.write("ncu --set full --import-source on --clock-control none, one launch (numbers under a profiler are not bench values)\n\n")
		f.write(f"grid {d.get('Grid Size')}  block {d.get('Block Size')}  cluster {d.get('launch__cluster_size', '-')}  "
				f"registers/thread {d.get('launch__registers_per_thread')}  dynamic smem/block {d.get('launch__shared_mem_per_block_dynamic')} {u.get('launch__shared_mem_per_block_dynamic', '')}\n")
		f.write(f"duration                         {dur:10.2f} us\n")
		f.write(f"dram__bytes_read.sum             {rd / 1e6:10.3f} MB\n")
		f.write(f"dram__bytes_write.sum            {wr / 1e6:10.3f} MB\n")
		f.write(f"traffic (read + write)           {(rd + wr) / 1e6:10.3f} MB  -> {(rd + wr) / 1e3 / dur if dur == dur and dur > 0 else float('nan'):8.1f} GB/s DRAM during this launch\n")
		f.write(f"IPC per SM (active cycles)       {num('sm__inst_executed.avg.per_cycle_active'):10.3f}\n")
		f.write(f"warp instructions executed       {num('smsp__inst_executed.sum'):10.0f}\n")
		f.write(f"active warps, % of peak          {num('sm__warps_active.avg.pct_of_peak_sustained_active'):10.2f}\n\n")
		f.write("warp stall samples (all warps; 'barrier' is mostly the idle warps of a block waiting for the colour's few active warps)\n")
		for v, k in stalls[:10]:
			f.write(f"  {k:24s} {int(v):7d}  {100.0 * v / total:5.1f}%\n")


def main() -> None:
	rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
	PROF.mkdir(exist_ok=True)
	launch_summary(rnd)
	for name, note in (("island", "many_pyramids, one block per bin"), ("scatter", "many_pyramids, single-pass partition for one block per bin"),
					   ("partition", "large_pyramid, two-phase partition for a cluster with owner lists"),
					   ("cluster", "large_pyramid, one 16-block cluster for the single island"),
					   ("grid", "joint_grid with B2GPU_LITE_JOINTS=0, grid-barrier kernel: with 256-byte joint records the island fits no cluster"),
					   ("cluster_joints", "joint_grid, one 16-block cluster: 19 800 plain revolute joints as 27-word records"),
					   ("assemble", "joint_grid, resident mode: the step's joint records from the table, the previous outputs and the uploaded runs"),
					   ("island_batch", "batch of 8192 small_pyramid worlds")):
		kernel_summary(rnd, name, note)
	for f in list(OUT.glob(f"{rnd}_bench_*.json")) + list(OUT.glob(f"{rnd}_e2e_trace.txt")):
		if f.stat().st_size > 0:
			shutil.copy(f, PROF / f.name)


if __name__ == "__main__":
	main()
