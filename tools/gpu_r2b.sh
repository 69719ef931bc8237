#!/bin/bash
# Experiment call: parity tests on the shipped build, then knob sweeps of the fused kernel on the tuning build.
tag=${1:-r2b}
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
done
cat $out/bench_k1.log
export OKP_TUNING_LIBRARY=$PWD/object_keypoints_b200/libokp_tuning.so
run() { echo "== $*" >> $out/sweep.log; env "$@" timeout 200 python tools/bench_k1.py $SHAPE $FRAMES 10 f32 fusedonly $EXTRA 2>&1 | grep -v Warning >> $out/sweep.log; }
SHAPE=180x320 FRAMES=4096
EXTRA="nocam" run A=0
EXTRA="lean" run A=0
EXTRA="" run OKP_STRIP_SMEM_KB=220 OKP_STRIP_THREADS=512 OKP_STREAM_EPILOGUE_WARPS=3
EXTRA="" run OKP_STRIP_SMEM_KB=220 OKP_STRIP_THREADS=512 OKP_STREAM_EPILOGUE_WARPS=2
EXTRA="" run OKP_STRIP_STAGES=3
EXTRA="" run OKP_STRIP_STAGES=5
SHAPE=64x64 FRAMES=32768
EXTRA="lean nocam" run A=0
EXTRA="lean" run OKP_STREAM_EPILOGUE_WARPS=3
EXTRA="lean" run OKP_STREAM_EPILOGUE_WARPS=3 OKP_STRIP_THREADS=192
EXTRA="lean" run OKP_STRIP_SMEM_KB=220 OKP_STRIP_THREADS=512 OKP_STREAM_EPILOGUE_WARPS=4
EXTRA="lean" run OKP_STRIP_SMEM_KB=220 OKP_STRIP_THREADS=448 OKP_STREAM_EPILOGUE_WARPS=5
EXTRA="lean" run OKP_STRIP_SMEM_KB=72 OKP_STRIP_THREADS=128 OKP_STREAM_EPILOGUE_WARPS=2
EXTRA="lean" run OKP_STRIP_SMEM_KB=72 OKP_STRIP_THREADS=160 OKP_STREAM_EPILOGUE_WARPS=1
cat $out/sweep.log
ls -la $out
