#!/bin/bash
# 64x64: thread cap variants, fresh --set full captures of the stream kernel (f32, bf16) with source counters.
tag=${1:-r2h}
out=gpurun_out/$tag
mkdir -p $out
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
done
grep -v Warn $out/bench_k1.log
export OKP_TUNING_LIBRARY=$PWD/object_keypoints_b200/libokp_tuning.so
run() { echo "== $*" >> $out/sweep.log; env "$@" timeout 200 python tools/bench_k1.py $SHAPE $FRAMES 10 $DT $EXTRA 2>&1 | grep -v Warning >> $out/sweep.log; }
SHAPE=64x64 FRAMES=32768 EXTRA="lean"
for DT in f32 bf16; do
run OKP_STRIP_THREADS=224
run OKP_STRIP_THREADS=224 OKP_STRIP_STAGES=3
run OKP_STRIP_THREADS=192 OKP_STREAM_EPILOGUE_WARPS=3
run OKP_STRIP_THREADS=160 OKP_STREAM_EPILOGUE_WARPS=2 OKP_STRIP_SMEM_KB=72
done
cat $out/sweep.log | cut -c1-400
unset OKP_TUNING_LIBRARY
for dt in f32 bf16; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:okp_peaks_stream -s 2 -c 1 -o $out/prof_64_$dt \
    python tools/bench_k1.py 64x64 32768 2 $dt lean > $out/ncu_64_$dt.log 2>&1
ncu -i $out/prof_64_$dt.ncu-rep --page source --csv > $out/src64_$dt.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:okp_group -s 2 -c 1 -o $out/prof_group_64 \
    python tools/bench_k1.py 64x64 32768 2 f32 lean > $out/ncu_group.log 2>&1
ncu -i $out/prof_group_64.ncu-rep --page source --csv > $out/src_group64.csv 2>/dev/null
ls -la $out
