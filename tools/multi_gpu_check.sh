#!/bin/bash
# Multi-GPU checks on an N-GPU box: bench.py under torchrun, and the gcn10 executable's per-GPU block queue.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== gcn10 executable on $N GPUs"
timeout 600 python tools/program_run.py --gpus $N --blocks 8 --size 9000 2>&1 | tail -15
