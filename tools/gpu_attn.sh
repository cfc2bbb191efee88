#!/bin/bash
# attention kernel check + timing: tools/gpu_attn.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
timeout 600 python tools/kbench.py attn_ln > gpurun_out/${TAG}_kbench_attn.txt 2>&1; cat gpurun_out/${TAG}_kbench_attn.txt
