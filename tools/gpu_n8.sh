#!/bin/bash
# One 8-GPU box: the strong-scaling curve of BASELINE config 2 at N = 1, 2, 4, 8 (same box, back to back),
# config 3 at 8 GPUs and the config-5 sweep at 8 GPUs.   gpurun --gpus 8 -- bash tools/gpu_n8.sh <tag>
tag=${1:-r2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/${tag}_n8_gpus.txt
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
$TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${tag}_scale_n2.json 2> gpurun_out/${tag}_scale_n2.err
$TR --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/${tag}_scale_n4.json 2> gpurun_out/${tag}_scale_n4.err
$TR --nproc-per-node 8 --master-port 29603 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/${tag}_scale_n8.json 2> gpurun_out/${tag}_scale_n8.err
$TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --workload config3 --steps 3 --warmup 3 > gpurun_out/${tag}_config3_n8.json 2> gpurun_out/${tag}_config3_n8.err
$TR --nproc-per-node 8 --master-port 29605 tools/sweep.py > gpurun_out/${tag}_sweep_n8.md 2> gpurun_out/${tag}_sweep_n8.err
tail -c 600 gpurun_out/${tag}_scale_n8.json
