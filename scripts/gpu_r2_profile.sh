#!/bin/bash
# Numbers of record for a round: [GPU suite,] bench line, ncu launch list of the same bench command, ncu --set full captures of the stage
# kernel, the Poisson / projection kernels and the high-order stage kernel. The .ncu-rep files are reduced to their raw / source CSV pages
# on the box (gpurun brings back at most 64 MiB).      gpurun --timeout 2400 -- 'bash scripts/gpu_r2_profile.sh TAG [tests]'
TAG=${1:-r2k}
OUT=gpurun_out; mkdir -p $OUT
if [[ "$2" == *tests* ]]; then
  (time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider) > $OUT/${TAG}_pytest.log 2>&1; tail -n 6 $OUT/${TAG}_pytest.log
fi
(time timeout 600 python bench.py) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo bench rc=$?; head -c 1500 $OUT/${TAG}_bench.json; echo
timeout 360 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ensemble > $OUT/${TAG}_ncu_bench.log 2>&1; echo launches rc=$?
reduce() {   # name: raw page always, SASS source page gzipped, report deleted
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/$1_source_sass.csv.gz
  rm -f $OUT/$1.ncu-rep
}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -c 3 -f -o $OUT/${TAG}_stage python scripts/profile_run.py 512 2 1 > $OUT/${TAG}_ncu_stage.log 2>&1; echo stage rc=$?; reduce ${TAG}_stage
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"poisson|fft_x|thomas|project" -s 12 -c 6 -f -o $OUT/${TAG}_poisson python scripts/profile_run.py 512 2 1 > $OUT/${TAG}_ncu_poisson.log 2>&1; echo poisson rc=$?; reduce ${TAG}_poisson
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stage_hi|specific" -s 2 -c 3 -f -o $OUT/${TAG}_stage_hi python scripts/hi_profile_run.py 256 2 > $OUT/${TAG}_ncu_hi.log 2>&1; echo hi rc=$?; reduce ${TAG}_stage_hi
du -sh $OUT; ls -la $OUT | tail -n 12
