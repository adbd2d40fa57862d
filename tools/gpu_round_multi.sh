#!/usr/bin/env bash
# tools/gpu_round_multi.sh <tag> <ngpus> — multi-GPU round (run under `gpurun --gpus N`): parity of both distributed
# modes against the oracle, then the contract bench at N GPUs (C2 weak scaling with broadcast, C3 strong scaling with
# the all-to-all shuffle).  Everything lands in gpurun_out/<tag>/.
set -u
TAG=${1:-m1}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > "$OUT/gpus.txt" 2>&1
timeout 400 $TR --master-port 29511 tests/dist_gpu_check.py --rows 4000000 > "$OUT/dist_check.log" 2>&1
echo "dist_check exit $?" >> "$OUT/dist_check.log"
tail -n 4 "$OUT/dist_check.log"
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/bench_c2_n$N.json" 2> "$OUT/bench_c2_n$N.err"
echo "bench c2 exit $?"; cat "$OUT/bench_c2_n$N.json"
[ -n "${SKIP_C3:-}" ] || timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --config C3 > "$OUT/bench_c3_n$N.json" 2> "$OUT/bench_c3_n$N.err"
echo "bench c3 exit $?"; cat "$OUT/bench_c3_n$N.json"
[ -n "${SKIP_REF:-}" ] || timeout 300 $TR --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > "$OUT/bench_ref_n$N.json" 2> "$OUT/bench_ref_n$N.err"
echo "bench ref exit $?"; cat "$OUT/bench_ref_n$N.json"
ls -la "$OUT"
