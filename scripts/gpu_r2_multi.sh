#!/bin/bash
# Multi-GPU leg: gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_r2_multi.sh TAG N [SIZE] [legs]'   legs: check tests bench ab
TAG=${1:-r2m}; N=${2:-2}; SIZE=${3:-512}; LEGS=${4:-"check bench ab"}
OUT=gpurun_out; mkdir -p $OUT
if [[ $LEGS == *check* ]]; then
  # P-GPU == 1-GPU on a 64^3-type case first (peer-memory path): a broken exchange must not cost the bench legs their timeouts
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/multi_gpu_check.py 64 3 > $OUT/${TAG}_check.log 2>&1
  grep MULTI_GPU $OUT/${TAG}_check.log || { echo "multi-GPU check FAILED"; tail -n 25 $OUT/${TAG}_check.log; exit 1; }
fi
if [[ $LEGS == *tests* ]]; then
  ( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider ) > $OUT/${TAG}_pytest.log 2>&1
  tail -n 6 $OUT/${TAG}_pytest.log
fi
run() {  # label, env...
  local label=$1; shift
  ( env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --size $SIZE --steps 10 --warmup 3 --no-e2e ) > $OUT/${TAG}_bench_${label}.json 2> $OUT/${TAG}_bench_${label}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_${label}.json") if l.startswith("{")][-1])
    print("$label", "N=$N", round(d["value"], 1), "Mcell/s", round(d["ms_per_step"], 3), "ms", d["breakdown_ms_per_step"], d["checks"].get("vs_single_gpu"))
except Exception as e:
    print("$label failed", e); print(open("$OUT/${TAG}_bench_${label}.err").read()[-1500:])
PY
}
if [[ $LEGS == *bench* ]]; then run flags_overlap BZ_DUMMY=1; fi
if [[ $LEGS == *ab* ]]; then
  run direct_pull BZ_DIRECT_PULL=1
  run packed_serial BZ_NO_OVERLAP=1
fi
if [[ $LEGS == *chunks* ]]; then
  run dma_chunks1 BZ_FFT_Z_CHUNKS=1
  run dma_chunks2 BZ_FFT_Z_CHUNKS=2
  run dma_chunks8 BZ_FFT_Z_CHUNKS=8
  run pull_chunks1 BZ_DMA_TRANSPOSE=0
fi
if [[ $LEGS == *r1style* ]]; then
  run nccl_serial_direct BZ_NO_OVERLAP=1 BZ_NCCL_BARRIER=1 BZ_DIRECT_PULL=1
fi
