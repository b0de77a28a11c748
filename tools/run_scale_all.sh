#!/bin/bash
# bench.py at 1/2/4/8 GPUs back to back on one 8-GPU box (what the driver's scaling run does); results -> gpurun_out/.
# Run under: gpurun --gpus 8 --timeout 900 -- 'bash tools/run_scale_all.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline > $OUT/scale_n1.json 2> $OUT/scale_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > $OUT/scale_n$N.json 2> $OUT/scale_n$N.err
done
for N in 1 2 4 8; do wc -l < $OUT/scale_n$N.json; python tools/brief.py $OUT/scale_n$N.json; done
