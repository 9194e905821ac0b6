#!/bin/bash
# Scaling records as the driver takes them: gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_r2_scale.sh TAG N [1024]'
TAG=${1:-r2s}; N=${2:-2}; BIG=$3
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/multi_gpu_check.py 64 3 > $OUT/${TAG}_check.log 2>&1
grep MULTI_GPU $OUT/${TAG}_check.log || { echo "multi-GPU check FAILED"; tail -n 25 $OUT/${TAG}_check.log; exit 1; }
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 ) > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench N=$N rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N", round(d["value"], 1), "Mcell/s", round(d["ms_per_step"], 3), "ms", d["breakdown_ms_per_step"], "e2e", d["e2e"]["value"], d["checks"].get("vs_single_gpu"))
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_n$N.err").read()[-1500:])
PY
( timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > $OUT/${TAG}_bench_ref_n$N.json 2> $OUT/${TAG}_bench_ref_n$N.err
python -c "
import json; d=json.loads([l for l in open('$OUT/${TAG}_bench_ref_n$N.json') if l.startswith('{')][-1]); print('reference arm under torchrun:', d['value'], d['cpu_baseline']['cores'], 'cores')"
if [ -n "$BIG" ]; then
  ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --size $BIG --steps 10 --warmup 3 --no-e2e ) > $OUT/${TAG}_bench_n${N}_$BIG.json 2> $OUT/${TAG}_bench_n${N}_$BIG.err
  echo "bench N=$N size=$BIG rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_n${N}_$BIG.json") if l.startswith("{")][-1])
    print("N=$N $BIG^3", round(d["value"], 1), "Mcell/s", round(d["ms_per_step"], 3), "ms", d["breakdown_ms_per_step"], d["checks"])
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_n${N}_$BIG.err").read()[-1500:])
PY
fi
