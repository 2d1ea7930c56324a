import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "kernel_ms", round(d["ms_per_step"],4), "e2e_ms", round(d["e2e"]["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), d["roofline"].get("island_bins_blocks_per_bin"), {k:round(v*1000,1) for k,v in d.get("stage_ms_per_step",{}).items()})
    except Exception as e: print(f, "ERR", e)
