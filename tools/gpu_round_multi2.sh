#!/usr/bin/env bash
# tools/gpu_round_multi2.sh <tag> <ngpus> [steps...] — multi-GPU round 2 (run under `gpurun --gpus N`).
# steps: check bench_c3 bench_c2 bench_c4 bench_c5 ref
set -u
TAG=${1:-m1}
N=${2:-2}
shift 2 || true
STEPS=${*:-check bench_c3}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ " $STEPS " == *" $1 "* ]]; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > "$OUT/gpus.txt" 2>&1
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
if has check; then
  timeout 400 $TR --master-port 29511 tests/dist_gpu_check.py --rows 4000000 > "$OUT/dist_check.log" 2>&1
  echo "dist_check exit $?" >> "$OUT/dist_check.log"
  tail -n 4 "$OUT/dist_check.log" | cut -c1-3000
fi
if has bench; then  # the contract line at N GPUs: C3 strong scaling + C4 / C2 under other_configs
  timeout 900 $TR --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
  echo "bench exit $?"; cut -c1-1500 "$OUT/bench_n$N.json"; tail -3 "$OUT/bench_n$N.err"
fi
if has bench_c3; then
  timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --config C3 > "$OUT/bench_c3_n$N.json" 2> "$OUT/bench_c3_n$N.err"
  echo "bench c3 exit $?"; cat "$OUT/bench_c3_n$N.json"; tail -3 "$OUT/bench_c3_n$N.err"
fi
if has bench_c2; then
  timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --config C2 > "$OUT/bench_c2_n$N.json" 2> "$OUT/bench_c2_n$N.err"
  echo "bench c2 exit $?"; cat "$OUT/bench_c2_n$N.json"; tail -3 "$OUT/bench_c2_n$N.err"
fi
if has bench_c4; then
  timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 --config C4 > "$OUT/bench_c4_n$N.json" 2> "$OUT/bench_c4_n$N.err"
  echo "bench c4 exit $?"; cat "$OUT/bench_c4_n$N.json"; tail -3 "$OUT/bench_c4_n$N.err"
fi
if has bench_c5; then
  timeout 600 $TR --master-port 29516 bench.py --gpus $N --steps 5 --warmup 3 --config C5 > "$OUT/bench_c5_n$N.json" 2> "$OUT/bench_c5_n$N.err"
  echo "bench c5 exit $?"; cat "$OUT/bench_c5_n$N.json"; tail -3 "$OUT/bench_c5_n$N.err"
fi
if has ref; then
  timeout 300 $TR --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > "$OUT/bench_ref_n$N.json" 2> "$OUT/bench_ref_n$N.err"
  echo "bench ref exit $?"; cat "$OUT/bench_ref_n$N.json"
fi
ls -la "$OUT"
