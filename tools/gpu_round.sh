#!/usr/bin/env bash
# tools/gpu_round.sh <tag> [steps...] — one gpurun call's worth of work; everything lands in gpurun_out/<tag>/.
# steps: sanity tests quick bench ref launches ncu_c2 ncu_c3 ncu_general ubench2 sanitize   (default: all but ncu_general, sanitize)
set -u
TAG=${1:-r1}
shift || true
STEPS=${*:-sanity tests quick bench ref launches ncu_c2 ncu_c3}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export FJ_OUT=$OUT
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,memory.total,power.limit --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host_cores.txt"; free -g >> "$OUT/host_cores.txt"

if has sanity; then
  # the direct-address join synchronises CTAs by spinning: make sure it terminates and is right on a small case
  # before anything larger depends on it; if not, the rest of the round runs with that path switched off
  timeout 180 python tools/prof_case.py S radix mat --check --reps 2 --set dense_min_rows=1024 > "$OUT/sanity.log" 2>&1
  rc=$?
  echo "sanity exit $rc" >> "$OUT/sanity.log"
  if [ $rc -ne 0 ]; then export FJ_CFG_DENSE_MIN_ROWS=4000000000; echo "dense radix path DISABLED for this round" >> "$OUT/sanity.log"; fi
  timeout 180 compute-sanitizer --tool memcheck python tools/prof_case.py S radix mat --reps 1 --set dense_min_rows=1024 > "$OUT/memcheck_dense.log" 2>&1
  # the fused bitmap count meets at spinning grid barriers: same precaution (small and large bitmap variant)
  { timeout 120 python tools/prof_case.py C1 scalar count --check --reps 3 && timeout 120 python tools/prof_case.py S scalar count bloom --check --reps 3; } > "$OUT/sanity_fused.log" 2>&1
  rc=$?
  echo "sanity_fused exit $rc" >> "$OUT/sanity_fused.log"
  if [ $rc -ne 0 ]; then export FJ_CFG_DENSE_FUSED=0; echo "fused bitmap count DISABLED for this round" >> "$OUT/sanity_fused.log"; fi
  timeout 180 compute-sanitizer --tool memcheck python tools/prof_case.py S scalar count --reps 1 > "$OUT/memcheck_fused.log" 2>&1
  tail -n 4 "$OUT/sanity_fused.log" "$OUT/memcheck_fused.log"
  tail -n 3 "$OUT/sanity.log" "$OUT/memcheck_dense.log"
fi
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -5 "$OUT/pytest_gpu.log"
fi
if has quick; then
  timeout 900 python tools/quick_bench.py C1 C2 C4s C3 --reps 5 > "$OUT/quick_bench.jsonl" 2> "$OUT/quick_bench.err"
  cat "$OUT/quick_bench.jsonl"
fi
if has bench; then
  timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
  cat "$OUT/bench.json"
fi
if has ref; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  cat "$OUT/bench_ref.json"
fi
if has launches; then
  # launch list of the bench command itself (times under ncu are cold-cache/serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/launches_bench.log" 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches_c3.csv" \
    python tools/prof_case.py C3 radix mat --reps 2 > "$OUT/launches_c3.log" 2>&1
fi
if has ncu_c2; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_count_dense|k_probe" -s 2 -c 1 -f -o "$OUT/c2_probe" \
    python tools/prof_case.py C2 scalar count bloom --reps 4 > "$OUT/ncu_c2.log" 2>&1
fi
if has ncu_c3; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_scatter2|k_djoin|k_join" -s 3 -c 3 -f -o "$OUT/c3_radix" \
    python tools/prof_case.py C3 radix mat --reps 3 > "$OUT/ncu_c3.log" 2>&1
fi
if has ncu_general; then  # the general hash path of the same two workloads (dense key domain switched off)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_probe -s 2 -c 1 -f -o "$OUT/c2_probe_general" \
    python tools/prof_case.py C2 scalar count bloom --reps 4 --set dense=0 > "$OUT/ncu_c2_general.log" 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_scatter2|k_join" -s 5 -c 5 -f -o "$OUT/c3_radix_general" \
    python tools/prof_case.py C3 radix mat --reps 3 --set dense=0 > "$OUT/ncu_c3_general.log" 2>&1
fi
if has ubench2; then  # design inputs for DESIGN.md §8 item 1 (L2 vs DSMEM vs shared memory for the random accesses)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/ubench2.cu -o "$OUT/ubench2.bin" > "$OUT/ubench2.build.log" 2>&1 \
    && timeout 300 "$OUT/ubench2.bin" > "$OUT/ubench2.jsonl" 2> "$OUT/ubench2.err"
  cat "$OUT/ubench2.jsonl"
fi
if has sanitize; then
  timeout 900 compute-sanitizer --tool memcheck python tools/prof_case.py S scalar mat bloom --reps 1 > "$OUT/memcheck.log" 2>&1
  timeout 900 compute-sanitizer --tool racecheck python tools/prof_case.py S radix mat --reps 1 > "$OUT/racecheck.log" 2>&1
fi
ls -la "$OUT"
