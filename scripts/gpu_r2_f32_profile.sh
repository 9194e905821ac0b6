OUT=gpurun_out; mkdir -p $OUT; TAG=r2y
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -c 3 -f -o $OUT/${TAG}_f32_stage python scripts/profile_run.py 512 2 1 Float32 > $OUT/${TAG}_ncu_f32_stage.log 2>&1; echo rc=$?
ncu -i $OUT/${TAG}_f32_stage.ncu-rep --page raw --csv > $OUT/${TAG}_f32_stage_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_f32_stage.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $OUT/${TAG}_f32_stage_source_sass.csv.gz
rm -f $OUT/${TAG}_f32_stage.ncu-rep; ls -la $OUT | grep r2y; tail -3 $OUT/${TAG}_ncu_f32_stage.log
