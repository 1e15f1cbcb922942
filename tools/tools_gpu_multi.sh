#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "bench n=$N exit $?" >> gpurun_out/bench_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "ref n=$N exit $?" >> gpurun_out/bench_ref_n$N.log
grep -E '^\{|exit|Error|error' gpurun_out/bench_n$N.log | cut -c1-1500 | tail -5
grep -E '^\{|exit|Error|error' gpurun_out/bench_ref_n$N.log | cut -c1-800 | tail -3
