#!/bin/bash
TAG=${1:-rX}
mkdir -p gpurun_out
( timeout 300 python tools/kbench_attn_variants.py
for f in tools/experiments/lib/lib*.so; do ESMDIFF_LIB=$PWD/$f timeout 300 python tools/kbench_attn_variants.py; done ) > gpurun_out/${TAG}_attn_variants.txt 2>&1
cat gpurun_out/${TAG}_attn_variants.txt
