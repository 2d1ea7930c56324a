import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d["config"]["workload"], "value %.3g"%d["value"], "kernel_ms %.4f"%d["ms_per_step"], "e2e_ms %.4f"%d["e2e"]["ms_per_step"], "plan", d["roofline"].get("island_bins_blocks_per_bin"), "gb", d["roofline"].get("grid_barriers_per_step"), "cpu %.3g"%d.get("cpu_baseline",{}).get("value",float("nan")), "frac", d["roofline"]["frac"])
