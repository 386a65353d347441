#!/bin/bash
# One gpurun call: GPU tests (with prints), the default bench line, optional extras.  Outputs -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
tail -5 gpurun_out/bench.err
