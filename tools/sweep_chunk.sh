#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep_chunk.txt
for c in 48 96 144 192 288 384 1024; do
  echo "chunk $c: $(ROVITKAN_CHUNK_IMAGES=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"value": [0-9.]*' | head -2 | tr '\n' ' ')" | tee -a gpurun_out/sweep_chunk.txt
done
