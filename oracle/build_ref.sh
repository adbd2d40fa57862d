#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE, not product code.
# Compiles the UNMODIFIED reference translation unit (/root/reference/hash_join.cpp) where it
# lies into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):
#   oracle/_ref/plain/flash_join<ext>.so  — the reference as is (returns (count, seconds))
#   oracle/_ref/pairs/flash_join_pairs<ext>.so — same source (module renamed flash_join_pairs) with the three `return py::make_tuple(...)`
#       statements of the materialize drivers (hash_join.cpp:380, :444, :494) extended, by a sed
#       run on a temp copy, to also return result_keys/result_values (the reference computes the
#       pairs and drops them).  The temp copy is deleted; no reference source enters the repo.
# The only thing substituted is the allocator header: `mimalloc.h` is a one-line stub providing
# mi_version() (hash_join.cpp:596 is its single use), because the statically interposed mimalloc
# of the reference's CMake build crashes next to pandas/torch in this image.
# The reference's own CMake is NOT run (it needs the vendored mimalloc sub-build).
set -euo pipefail
REF=${REF_SRC:-/root/reference/hash_join.cpp}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) — keeping prebuilt $OUT" >&2
  exit 0
fi
PY=${PYTHON:-python}
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
INC=$($PY -m pybind11 --includes)
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT/plain" "$OUT/pairs" "$TMP/stub"
echo 'static inline int mi_version(void) { return 0; }' > "$TMP/stub/mimalloc.h"
# -march=x86-64-v2 (SSE4.2 for _mm_crc32_u64) + AVX2 when the build host has it; the GPU box may be a
# different CPU model than this container, so do not use -march=native for the artefact that travels.
CXXFLAGS="-O3 -msse4.2 -mavx2 -mtune=generic -std=c++17 -shared -fPIC -fvisibility=hidden -pthread"
g++ $CXXFLAGS $INC -I"$TMP/stub" "$REF" -o "$OUT/plain/flash_join$EXT"
# pairs variant: textual patch of the three materialize returns on a temp copy
sed -E -e 's/return py::make_tuple\(py::int_\(total_results\), core_duration_sec\);/return py::make_tuple(py::int_(total_results), core_duration_sec, result_keys, result_values);/' \
    -e 's/PYBIND11_MODULE\(flash_join, m\)/PYBIND11_MODULE(flash_join_pairs, m)/' \
    "$REF" > "$TMP/hash_join_pairs.cpp"
# count-path returns (:533, :566) have no result arrays in scope -> restore them
$PY - "$TMP/hash_join_pairs.cpp" <<'PYEOF'
import re, sys
p = sys.argv[1]
src = open(p).read().split('\n')
patched = 'return py::make_tuple(py::int_(total_results), core_duration_sec, result_keys, result_values);'
orig = 'return py::make_tuple(py::int_(total_results), core_duration_sec);'
out = []
n = 0
for i, line in enumerate(src):
    if patched in line:
        # keep the patch only where result_keys was declared earlier in the same function
        j = i
        ok = False
        while j > 0 and not src[j].startswith('py::tuple _hash_join') and not src[j].startswith('template'):
            if 'py::array_t<uint64_t> result_keys' in src[j]:
                ok = True
                break
            j -= 1
        if not ok:
            line = line.replace(patched, orig)
        else:
            n += 1
    out.append(line)
assert n == 3, f"expected 3 materialize returns, patched {n}"
open(p, 'w').write('\n'.join(out))
PYEOF
g++ $CXXFLAGS $INC -I"$TMP/stub" -I"$(dirname "$REF")" "$TMP/hash_join_pairs.cpp" -o "$OUT/pairs/flash_join_pairs$EXT"
# timing variant: the reference's vendored allocator (mimalloc, /root/reference/mimalloc) as a malloc-overriding shared
# library, built from its single-file source (src/static.c) where it lies; the reference's CMake itself is not run.
# bench.py's CPU worker — a numpy-only subprocess — runs the plain module above under LD_PRELOAD of this library, so
# every allocation of the process (std::vector / make_unique of the hot path included) goes through mimalloc, which is
# what the reference's CMakeLists.txt (MI_OVERRIDE / MI_MALLOC_OVERRIDE, :9-16, :23-28) arranges.  (Linking mimalloc
# into the module with hidden symbols crashed: libstdc++ freed the module's blocks with glibc's free.)  Never
# preloaded next to torch/pandas; parity always uses the plain build.
MI="$(dirname "$REF")/mimalloc"
if [ -f "$MI/src/static.c" ]; then
  mkdir -p "$OUT/mimalloc"
  rm -f "$OUT/mimalloc/flash_join$EXT"
  gcc -O3 -DNDEBUG -DMI_MALLOC_OVERRIDE -fPIC -shared -std=gnu11 -fno-builtin-malloc -ftls-model=initial-exec \
      -I"$MI/include" "$MI/src/static.c" -o "$OUT/mimalloc/libmimalloc_override.so" -lpthread
  echo "built: $OUT/mimalloc/libmimalloc_override.so"
fi
echo "built: $OUT/plain/flash_join$EXT $OUT/pairs/flash_join_pairs$EXT"
