#!/bin/bash
# ncu evidence: launch list of the bench command, full captures of K1 (180x320 f32, 64x64 f32 / bf16), the grouping kernel,
# launch list + full captures of the kernels beside the decode path.
tag=${1:-evidence}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:okp_ -c 80 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --e2e-frames 256 --e2e-steps 1 > $out/bench_under_ncu.log 2>&1
full() { # name regex command...
  name=$1; regex=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -o $out/$name "$@" > $out/$name.log 2>&1
  ncu -i $out/$name.ncu-rep --page raw --csv > $out/$name.raw.csv 2>/dev/null
  ncu -i $out/$name.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null
  python tools/ncu_hot.py /tmp/$name.src.csv 1.0 > $out/$name.hot.txt 2>&1
  rm -f $out/$name.ncu-rep
}
full k1_180_f32 okp_peaks_stream python tools/bench_k1.py 180x320 4096 2 f32
full k1_64_f32 okp_peaks_stream python tools/bench_k1.py 64x64 32768 2 f32 lean
full k1_64_bf16 okp_peaks_stream python tools/bench_k1.py 64x64 32768 2 bf16 lean
full group_180_f32 okp_group_kernel python tools/bench_k1.py 180x320 4096 2 f32
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:okp_ -c 200 --csv --log-file $out/secondary_launches.csv \
    python tools/bench_secondary.py > $out/secondary_under_ncu.log 2>&1
full correct_matches okp_correct_matches python tools/bench_secondary.py
full triangulate_robust okp_triangulate_robust python tools/bench_secondary.py
full rasterise okp_rasterise python tools/bench_secondary.py
rm -f $out/*.ncu-rep.tmp
ls -la $out
