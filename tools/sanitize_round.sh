#!/bin/bash
# compute-sanitizer pass of a round (run on the GPU box through gpurun): memcheck over the C-ABI tests and a few lockstep scenes
# through b2World_Step, racecheck over the chained resident steps (island kernel, clusters) -- output: gpurun_out/<round>_sanitizer.txt
R=${1:-r02}
O=gpurun_out/${R}_sanitizer.txt
mkdir -p gpurun_out
echo "compute-sanitizer on a B200 (gpurun), final round-2 build (deferred impulses and joint outputs, direct outputs, bin lists and plans kept, compact revolute records, host arrays on registered huge pages):" > $O
echo '--- memcheck: python -m pytest tests -m gpu -k "not lockstep and not large_batch and not many_worlds and not group_of_live and not nobody_looks and not shortcut"' >> $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests -m gpu -q -x -k "not lockstep and not large_batch and not many_worlds and not group_of_live and not nobody_looks and not shortcut" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -20 >> $O
echo '--- memcheck, through b2World_Step: tests/test_gpu_lockstep.py -k "small_pyramid or joint_zoo or mutator or contact_zoo" + tests/test_gpu_deferred.py -k "joint_reactions or snapshot"' >> $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_lockstep.py tests/test_gpu_deferred.py -m gpu -q -x -k "small_pyramid or joint_zoo or mutator or contact_zoo or joint_reactions or snapshot" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|Invalid|error" | head -20 >> $O
echo '--- racecheck: tests/test_gpu_resident.py -k "(chained and island) or clusters_run" + tests/test_gpu_capture.py -k on_clusters' >> $O
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_gpu_resident.py tests/test_gpu_capture.py -m gpu -q -x -k "(chained and island) or clusters_run or on_clusters" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|RACECHECK SUMMARY|hazard|error" | head -20 >> $O
cat $O
