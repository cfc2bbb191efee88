#!/bin/bash
# One GPU: smoke, the default bench line (with cpu baseline), the reference arm, ncu launch list and full-set capture (no sweep).
tag=${1:-r5}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 900 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attention_resident|gemm_bf16" --launch-skip 7 --launch-count 7 -o gpurun_out/${tag}_full python tools/prof_kernels.py > gpurun_out/${tag}_ncu_full.log 2>&1
tail -1 gpurun_out/${tag}_smoke.txt; tail -c 300 gpurun_out/${tag}_bench.json
