#!/bin/bash
# Knob sweep of the launch plans on the TUNING build (the shipped library reads no environment):
#   python -c "from object_keypoints_b200 import _lib; _lib.build(tuning=True, output='object_keypoints_b200/libokp_tuning.so')"
#   gpurun -- 'bash tools/gpu_sweep.sh <tag> <shape> <frames> <dtype> "KNOB=V KNOB=V" "KNOB=V" ...'
# Knobs (csrc/okp_peaks_strip.cuh, okp_peaks_stream.cuh, okp_peaks_tile.cuh, okp_api.cu): OKP_STRIP_STAGES, OKP_STRIP_THREADS,
# OKP_STRIP_SMEM_KB, OKP_STREAM_EPILOGUE_WARPS, OKP_STREAM_EDGE_ONLY, OKP_GROUP_LANES, OKP_PEAKS_TILE (+ OKP_TILE_*).
tag=${1:-sweep}; shape=${2:-180x320}; frames=${3:-4096}; dt=${4:-f32}; shift 4
out=gpurun_out/$tag
mkdir -p $out
export OKP_TUNING_LIBRARY=$PWD/object_keypoints_b200/libokp_tuning.so
extra=""; [ "$shape" = "64x64" ] && extra="lean"
echo "== shipped plan" >> $out/sweep.log
timeout 200 python tools/bench_k1.py $shape $frames 10 $dt $extra 2>&1 | grep -v Warning >> $out/sweep.log
for knobs in "$@"; do
  echo "== $knobs" >> $out/sweep.log
  env $knobs timeout 200 python tools/bench_k1.py $shape $frames 10 $dt $extra 2>&1 | grep -v Warning >> $out/sweep.log
done
sed 's/env={[^}]*}//' $out/sweep.log | cut -c1-220
