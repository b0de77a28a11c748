#!/bin/bash
# Round-end evidence on one B200: GPU tests, the default bench line, the ncu launch list of the same command,
# one --set full capture of the depth 0-2 kernels, DRAM traffic of the traversal launches.  Output -> gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $OUT/bench_final.json 2> $OUT/bench_final.err
python tools/brief.py $OUT/bench_final.json
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_final.csv $BENCH > /dev/null 2> $OUT/ncu_launches.err
ncu --set full --clock-control none --import-source on -k 'regex:k_trace_dual|k_shade|k_extend_primary' -s 16 -c 5 -o $OUT/full_final $BENCH > /dev/null 2> $OUT/ncu_full.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:k_extend|k_trace_dual|k_connect' -s 9 -c 9 --csv --log-file $OUT/traffic_final.csv $BENCH > /dev/null 2> $OUT/ncu_traffic.err
ls -la $OUT/full_final.ncu-rep $OUT/launches_final.csv $OUT/traffic_final.csv
