#!/usr/bin/env bash
# tools/gpu_round2.sh <tag> [steps...] — one gpurun call's worth of round-2 work; everything lands in gpurun_out/<tag>/.
# steps: sanity sanitize tests16 tests quick quickfull bench ref launches ncu_c3 ncu_general ubench2
set -u
TAG=${1:-r2}
shift || true
STEPS=${*:-sanity sanitize tests16 quick launches ncu_c3}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,memory.total,power.limit --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host_cores.txt"; free -g >> "$OUT/host_cores.txt"

if has sanity; then
  # new kernels first on small cases, each under its own timeout: a hang must not eat the call
  for cfg in T S; do
    for mode in mat count; do
      timeout 120 python tools/prof_case.py $cfg radix $mode --check --reps 2 --set dense_min_rows=1024 >> "$OUT/sanity.log" 2>&1
      echo "sanity $cfg $mode exit $?" >> "$OUT/sanity.log"
    done
  done
  tail -n 30 "$OUT/sanity.log"
fi
if has sanitize; then
  timeout 300 compute-sanitizer --tool memcheck python tools/prof_case.py S radix mat --reps 1 --set dense_min_rows=1024 > "$OUT/memcheck_dense16.log" 2>&1
  echo "exit $?" >> "$OUT/memcheck_dense16.log"
  timeout 300 compute-sanitizer --tool racecheck python tools/prof_case.py T radix mat --reps 1 --set dense_min_rows=1024 > "$OUT/racecheck_dense16.log" 2>&1
  echo "exit $?" >> "$OUT/racecheck_dense16.log"
  tail -n 6 "$OUT/memcheck_dense16.log" "$OUT/racecheck_dense16.log"
fi
if has sanitize_general; then  # the general-path kernels under memcheck + racecheck
  timeout 600 compute-sanitizer --tool memcheck python tools/prof_case.py S radix mat --reps 1 --set dense=0 > "$OUT/memcheck_radix_general.log" 2>&1
  echo "exit $?" >> "$OUT/memcheck_radix_general.log"
  timeout 600 compute-sanitizer --tool racecheck python tools/prof_case.py T radix mat --reps 1 --set dense=0 > "$OUT/racecheck_radix_general.log" 2>&1
  echo "exit $?" >> "$OUT/racecheck_radix_general.log"
  timeout 600 compute-sanitizer --tool memcheck python tools/prof_case.py S scalar mat bloom --reps 1 --set dense=0 > "$OUT/memcheck_scalar_general.log" 2>&1
  echo "exit $?" >> "$OUT/memcheck_scalar_general.log"
  timeout 600 compute-sanitizer --tool racecheck python tools/prof_case.py T scalar mat bloom --reps 1 --set dense=0 > "$OUT/racecheck_scalar_general.log" 2>&1
  echo "exit $?" >> "$OUT/racecheck_scalar_general.log"
  tail -n 4 "$OUT"/memcheck_*_general.log "$OUT"/racecheck_*_general.log
fi
if has tests16; then
  timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method thread -k "dense16 or smoke or adaptive" > "$OUT/pytest_dense16.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_dense16.log"
  tail -15 "$OUT/pytest_dense16.log"
fi
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -8 "$OUT/pytest_gpu.log"
fi
if has quick; then
  timeout 600 python tools/quick_bench.py C3 --reps 5 > "$OUT/quick_bench.jsonl" 2> "$OUT/quick_bench.err"
  cat "$OUT/quick_bench.jsonl"; tail -3 "$OUT/quick_bench.err"
fi
if has quickfull; then
  timeout 900 python tools/quick_bench.py C1 C2 C4s C3 --reps 5 --full > "$OUT/quick_bench_full.jsonl" 2> "$OUT/quick_bench_full.err"
  cat "$OUT/quick_bench_full.jsonl"
fi
if has bench; then
  timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
  cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
fi
if has ref; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  cat "$OUT/bench_ref.json"
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches_c3.csv" \
    python tools/prof_case.py C3 radix mat --reps 3 > "$OUT/launches_c3.log" 2>&1
  grep -E "k_part|k_sjoin|k_prepare" "$OUT/launches_c3.csv" | tail -8
fi
if has ncu_c3; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_part|k_sjoin" -s 3 -c 3 -f -o "$OUT/c3_dense16" \
    python tools/prof_case.py C3 radix mat --reps 3 > "$OUT/ncu_c3.log" 2>&1
  tail -3 "$OUT/ncu_c3.log"
fi
if has ncu_general; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_scatter2|k_join" -s 5 -c 5 -f -o "$OUT/c3_radix_general" \
    python tools/prof_case.py C3 radix mat --reps 3 --set dense=0 > "$OUT/ncu_c3_general.log" 2>&1
fi
if has sweep_adaptive; then
  timeout 900 python tools/sweep_adaptive.py > "$OUT/sweep_adaptive.jsonl" 2> "$OUT/sweep_adaptive.err"
  cat "$OUT/sweep_adaptive.jsonl"; tail -3 "$OUT/sweep_adaptive.err"
fi
if has exp_selectivity; then
  timeout 600 python tools/exp_selectivity.py > "$OUT/exp_selectivity.jsonl" 2> "$OUT/exp_selectivity.err"
  cat "$OUT/exp_selectivity.jsonl"; tail -3 "$OUT/exp_selectivity.err"
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
  echo "smoke exit $?" >> "$OUT/smoke.log"; tail -3 "$OUT/smoke.log"
fi
if has ubench2; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/ubench2.cu -o "$OUT/ubench2.bin" > "$OUT/ubench2.build.log" 2>&1 \
    && timeout 300 "$OUT/ubench2.bin" > "$OUT/ubench2.jsonl" 2> "$OUT/ubench2.err"
  cat "$OUT/ubench2.jsonl"
fi
if has ubench3; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench3.cu -o "$OUT/ubench3.bin" > "$OUT/ubench3.build.log" 2>&1 \
    && timeout 120 "$OUT/ubench3.bin" > "$OUT/ubench3.jsonl" 2> "$OUT/ubench3.err"
  cat "$OUT/ubench3.jsonl"
fi
ls -la "$OUT"
