#!/bin/bash
# A/B of library variants on one B200: tools/ab.sh <tag> <variant> ... (variant "default" = cadrays_b200/libcadrays_b200.so,
# otherwise cadrays_b200/libcadrays_b200_<variant>.so built with `python -m cadrays_b200.build -D... --out=...`).
# One quick bench line per variant -> gpurun_out/<tag>_<variant>.json, summarised by tools/brief.py.
TAG=$1; shift
mkdir -p gpurun_out
ARGS=${AB_ARGS:---steps 6 --warmup 3 --no-cpu-baseline --no-extras}
for v in "$@"; do
  if [ "$v" = default ]; then unset CADRAYS_B200_LIB; else export CADRAYS_B200_LIB=$PWD/cadrays_b200/libcadrays_b200_$v.so; fi
  python bench.py $ARGS > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python tools/brief.py gpurun_out/${TAG}_$v.json
done
