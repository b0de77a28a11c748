#!/bin/bash
# Round-2 evidence on one B200 (gpurun): GPU tests, the default bench line, the memory probe sweep, the ncu launch
# list of the bench command with the counters the issue / L1 / L2 / DRAM fractions are computed from (C2 and C5
# flattened), one --set full capture of the depth 0-2 kernels.  Output -> gpurun_out/<tag>_*.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc >> $OUT/${TAG}_smi.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  python -m pytest tests -m gpu -x -q --durations=25 > $OUT/${TAG}_pytest.txt 2>&1
  tail -40 $OUT/${TAG}_pytest.txt
fi
env -u CRT_PIPELINE python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python tools/brief.py $OUT/${TAG}_bench.json
python -m cadrays_b200.probe > $OUT/${TAG}_mem_probe.json 2> $OUT/${TAG}_mem_probe.err
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active
# CRT_PIPELINE=0: a wave unsplit on one stream, so that a step is one run of launches in the list (ncu serialises kernels anyway)
export CRT_PIPELINE=0
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
ncu --metrics $M --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > /dev/null 2> $OUT/${TAG}_ncu_launches.err
ncu --metrics $M --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches_c5flat.csv $BENCH --workload instanced_flat > /dev/null 2> $OUT/${TAG}_ncu_launches_c5.err
if [ "${SKIP_FULL:-0}" != "1" ]; then
  ncu --set full --clock-control none --import-source on -k 'regex:k_trace_dual|k_shade|k_extend_primary' -s 16 -c 5 -o $OUT/${TAG}_full $BENCH > /dev/null 2> $OUT/${TAG}_ncu_full.err
fi
ls -la $OUT | grep ${TAG}
