#!/bin/bash
# parity tests + one bench line: tools/gpu_quick.sh TAG [extra env assignments for a second bench line]
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
if [ -n "$2" ]; then env $2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_alt.json 2> gpurun_out/${TAG}_bench_alt.err; cat gpurun_out/${TAG}_bench_alt.json; fi
