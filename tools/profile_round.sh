#!/bin/bash
# Profiling pass of one round (run on the GPU box through gpurun): bench lines of every BASELINE.json config, the ncu
# launch list of the default bench command and one --set full capture per kernel.  Outputs go to gpurun_out/.
R=${1:-r01}
O=gpurun_out
mkdir -p $O
Q="--no-batch --no-cpu-baseline"
timeout 400 python bench.py > $O/${R}_bench_default.json 2> $O/${R}_bench_default.err
for w in large_pyramid joint_grid rain tumbler; do
  timeout 300 python bench.py --workload $w --steps 60 --warmup 10 --no-batch 2>/dev/null | tail -1 > $O/${R}_bench_$w.json
done
B2GPU_TRACE=1 timeout 200 python bench.py --steps 8 --warmup 20 $Q 2>&1 >/dev/null | grep b2gpu | tail -4 > $O/${R}_e2e_trace.txt
timeout 300 python bench.py --workload batch --steps 10 --warmup 3 2>/dev/null | tail -1 > $O/${R}_bench_batch.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/${R}_bench_reference_arm.json
# the launch list of the default bench command (every launch of the timed regions; the scene part of the line)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_default.csv python bench.py --steps 3 --warmup 3 $Q > $O/${R}_ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:b2gIslandKernel -c 1 -s 12 -o $O/${R}_island -f python bench.py --steps 3 --warmup 3 $Q > $O/${R}_ncu_island.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:b2gScatterKernel -c 1 -s 12 -o $O/${R}_scatter -f python bench.py --steps 3 --warmup 3 $Q > $O/${R}_ncu_scatter.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:b2gPartitionKernel -c 1 -s 8 -o $O/${R}_partition -f python bench.py --workload large_pyramid --steps 3 --warmup 3 $Q > $O/${R}_ncu_partition.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:b2gClusterIslandKernel -c 1 -s 8 -o $O/${R}_cluster -f python bench.py --workload large_pyramid --steps 3 --warmup 3 $Q > $O/${R}_ncu_cluster.log 2>&1
B2GPU_LITE_JOINTS=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:b2gStepKernel -c 1 -s 8 -o $O/${R}_grid -f python bench.py --workload joint_grid --steps 3 --warmup 3 $Q > $O/${R}_ncu_grid.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:b2gClusterIslandKernel -c 1 -s 8 -o $O/${R}_cluster_joints -f python bench.py --workload joint_grid --steps 3 --warmup 3 $Q > $O/${R}_ncu_cluster_joints.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:b2gAssembleJointsKernel -c 1 -s 8 -o $O/${R}_assemble -f python bench.py --workload joint_grid --steps 3 --warmup 3 $Q > $O/${R}_ncu_assemble.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:b2gIslandKernel -c 1 -o $O/${R}_island_batch -f python bench.py --workload batch --steps 2 --warmup 3 > $O/${R}_ncu_island_batch.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw-metrics page of every capture, and only the island kernel's report itself
for n in scatter partition cluster cluster_joints grid assemble island_batch island; do
  [ -f $O/${R}_$n.ncu-rep ] && ncu -i $O/${R}_$n.ncu-rep --page raw --csv > $O/${R}_$n.rawpage.csv 2>/dev/null
  [ $n != island ] && rm -f $O/${R}_$n.ncu-rep
done
ls -la $O | tail -30
