#!/bin/bash
# counters of the first k_trace_dual launches (depth 1, 2) of a bench step for library variants: tools/ncu_dual.sh <tag> <variant> ...
TAG=$1; shift
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for v in "$@"; do
  if [ "$v" = default ]; then unset CADRAYS_B200_LIB; else export CADRAYS_B200_LIB=$PWD/cadrays_b200/libcadrays_b200_$v.so; fi
  ncu --metrics $M --clock-control none -k regex:"${NCU_K:-k_trace_dual}" -c ${NCU_C:-2} --csv --log-file gpurun_out/${TAG}_$v.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras ${NCU_BENCH_ARGS:-} > /dev/null 2> gpurun_out/${TAG}_$v.ncu.err
  python tools/launch_table.py gpurun_out/${TAG}_$v.csv | sed "s/^/$v /"
done
