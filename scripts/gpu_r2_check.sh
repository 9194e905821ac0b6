#!/bin/bash
# Round-2 GPU-box call: parity suite with the measured errors logged, library variants A/B, bench line + reference arm.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_check.sh TAG [legs]'      legs: any of  tests variants bench ncu  (default: all but ncu)
TAG=${1:-r2}
LEGS=${2:-"tests variants bench"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [[ $LEGS == *tests* ]]; then
  rm -f $OUT/${TAG}_parity_errors.txt
  ( time BZ_EXPERIMENTAL_WENO_ORDER=1 BZ_PARITY_REPORT=$PWD/$OUT/${TAG}_parity_errors.txt timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider ) > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -n 15 $OUT/${TAG}_pytest.log
fi
if [[ $LEGS == *variants* ]] && ls breeze.jl_b200/csrc/variants/*.so >/dev/null 2>&1; then
  timeout 420 python scripts/variant_bench.py $(ls breeze.jl_b200/csrc/variants/*.so) --steps 5 > $OUT/${TAG}_variants.log 2>&1
  cat $OUT/${TAG}_variants.log | tail -n 8
fi
if [[ $LEGS == *bench* ]]; then
  ( time timeout 600 python bench.py ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  echo "bench rc=$?"; tail -c 2500 $OUT/${TAG}_bench.json; tail -n 5 $OUT/${TAG}_bench.err
  ( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
fi
if [[ $LEGS == *ncu* ]]; then
  timeout 360 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
  echo "ncu launch list rc=$?"
  timeout 360 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -c 3 -f -o $OUT/${TAG}_stage python scripts/profile_run.py 512 2 1 > $OUT/${TAG}_ncu_stage.log 2>&1
  echo "ncu stage rc=$?"
fi
ls -la $OUT | tail -n 12
