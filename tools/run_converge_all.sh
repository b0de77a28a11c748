#!/bin/bash
# C4 (4K, 4096 spp) at 1/2/4/8 GPUs and C3 (depth 12, 1080p) at 1 and 8 GPUs on one 8-GPU box; results -> gpurun_out/.
# Run under: gpurun --gpus 8 --timeout 1200 -- 'bash tools/run_converge_all.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
SPP=${SPP:-4096}
python tools/converge.py --config c4 --spp $SPP --save /tmp/c4_n1.npy > $OUT/converge_c4_n1.json 2> $OUT/converge_c4_n1.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29540 + N)) tools/converge.py --config c4 --spp $SPP --compare /tmp/c4_n1.npy \
    > $OUT/converge_c4_n$N.json 2> $OUT/converge_c4_n$N.err
done
python tools/converge.py --config c3 --spp $SPP --save /tmp/c3_n1.npy > $OUT/converge_c3_n1.json 2> $OUT/converge_c3_n1.err
$TR --nproc-per-node 8 --master-port 29561 tools/converge.py --config c3 --spp $SPP --compare /tmp/c3_n1.npy \
  > $OUT/converge_c3_n8.json 2> $OUT/converge_c3_n8.err
# progressive variant: an all-reduce every 64 samples per rank
$TR --nproc-per-node 8 --master-port 29562 tools/converge.py --config c4 --spp $SPP --interval 64 --compare /tmp/c4_n1.npy \
  > $OUT/converge_c4_n8_interval64.json 2> $OUT/converge_c4_n8_interval64.err
# the bench at 8 GPUs: stdout must be exactly one JSON line
$TR --nproc-per-node 8 --master-port 29563 bench.py --gpus 8 --steps 20 --warmup 3 > $OUT/bench_n8_stdout.txt 2> $OUT/bench_n8.err
wc -l $OUT/bench_n8_stdout.txt
cat $OUT/converge_*.json
