#!/bin/bash
# Runs on the GPU box (gpurun): parity tests, the bench arms and the two ncu passes behind profiles/<tag>_*.
# Outputs land in gpurun_out/; tools/make_profile_summary.py turns them into the tracked summaries.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --flags 4 --no-cpu-baseline > gpurun_out/bench_fma.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
# launch list: every kernel of two serial passes of 16 frames
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --flags 16 > gpurun_out/ncu_launch.log 2>&1
# full capture of the 20 pyramid launches of one pass (one image group, so each launch covers all 16 frames)
SIFT_GPU_PYR_GROUPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_" -s 20 -c 20 -f \
    -o gpurun_out/pyramid_full python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --flags 16 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest_gpu.txt
tail -c 600 gpurun_out/bench.json
ls -la gpurun_out/
