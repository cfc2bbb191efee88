#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list of the bench command, and one
# `ncu --set full` capture of each hot kernel at the config-2 shape.  Usage: tools/gpu_round.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_bf16|attention_|layernorm_f32|qk_layernorm|embed_kernel|sample_rows|time_embed|fill_i64' -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# skip the set-up GEMM and the first two iterations of tools/ncu_targets.py (7 launches each), capture the third
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_tn_kernel|attention_resident_kernel|qk_layernorm_rope_kernel|sample_rows_kernel' --launch-skip 15 -c 7 -f -o gpurun_out/${TAG}_full \
    python tools/ncu_targets.py 100 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
