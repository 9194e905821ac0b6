#!/bin/bash
# One GPU-box call in priority order: GPU parity suite, A/B of library variants, bench line, ncu launch list, ncu full capture of the
# stage kernel. Every leg has its own timeout; outputs land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round_check.sh TAG'
TAG=${1:-chk}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
( time timeout 480 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
if ls breeze.jl_b200/csrc/variants/*.so >/dev/null 2>&1; then
  timeout 300 python scripts/variant_bench.py $(ls breeze.jl_b200/csrc/variants/*.so) --steps 5 > $OUT/${TAG}_variants.log 2>&1
  cat $OUT/${TAG}_variants.log | tail -5
fi
( time timeout 480 python bench.py ) > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 1500 $OUT/${TAG}_bench.json
timeout 360 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
timeout 360 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -c 3 -f -o $OUT/${TAG}_stage python scripts/profile_run.py 512 2 1 > $OUT/${TAG}_ncu_stage.log 2>&1
echo "ncu stage rc=$?"
ls -la $OUT | tail -15
