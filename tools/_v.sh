for v in 0 1; do
CRT_SHADOW0_LOCKSTEP=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/sl$v.json 2> gpurun_out/sl$v.err; python tools/brief.py gpurun_out/sl$v.json
done
