#!/bin/bash
# BASELINE configs[4]: ROI-head microbench sweep, 256-8192 proposals/img x 20/80 classes x 512-d text embeddings, 1 GPU,
# inference step; one bench line per point into gpurun_out/sweep.jsonl
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for K in 20 80; do
  for P in 256 512 1024 2048 4096 8192; do
    timeout 300 python bench.py --mode infer --props $P --classes $K --images-per-gpu 4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras \
      2> gpurun_out/sweep_err.log | tail -n 1 >> gpurun_out/sweep.jsonl || echo "{\"failed\": [$P, $K]}" >> gpurun_out/sweep.jsonl
  done
done
wc -l gpurun_out/sweep.jsonl
