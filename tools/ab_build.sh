#!/bin/bash
# Build libgpsb_cuda.so with extra nvcc flags and park it as tools/bin/ab/<name>.so (A/B timing of kernel variants on one
# GPU box: tools/ab_run.sh).  Usage: bash tools/ab_build.sh <name> [nvcc flags ...]; leaves the default build in lib/.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/bin/ab
GPSB_NVCC_EXTRA="$*" python -c "from stm32f4_sdr_gps_b200 import build as b; b.build_cuda(force=True)" >/dev/null
cp stm32f4_sdr_gps_b200/lib/libgpsb_cuda.so tools/bin/ab/$name.so
echo "tools/bin/ab/$name.so  ($*)"
