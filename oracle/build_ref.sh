#!/usr/bin/env bash
# Builds the UNMODIFIED reference Raven (read-only tree at $RAVEN_SRC, default /root/reference/src)
# into oracle/_ref/:   Raven.exe (CLI; CPU baseline)  and  libravenref.so (-D_BMI_LIBRARY_, all C++
# symbols exported; linked by oracle/ref_dump.cpp for in-memory FP64 parity dumps).
# TEST INFRASTRUCTURE ONLY - nothing under oracle/ is on the product path.
# Recipe follows SURVEY.md App. E (g++ -std=c++11 -O2, no NetCDF, no lp_solve). No reference
# source is copied: objects are compiled from where the sources lie; outputs go to oracle/_ref only.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${RAVEN_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then echo "build_ref: $SRC absent - keeping prebuilt $OUT" >&2; exit 0; fi
mkdir -p "$OUT/obj_exe" "$OUT/obj_lib"
NJ="${NJ:-$(nproc)}"
ls "$SRC"/*.cpp | xargs -P"$NJ" -I{} sh -c \
  'o='"$OUT"'/obj_exe/$(basename {} .cpp).o; [ "$o" -nt {} ] || g++ -std=c++11 -O2 -w -c {} -o "$o"'
g++ -o "$OUT/Raven.exe" "$OUT"/obj_exe/*.o
ls "$SRC"/*.cpp | xargs -P"$NJ" -I{} sh -c \
  'o='"$OUT"'/obj_lib/$(basename {} .cpp).o; [ "$o" -nt {} ] || g++ -std=c++11 -O2 -w -fPIC -D_BMI_LIBRARY_ -c {} -o "$o"'
g++ -shared -o "$OUT/libravenref.so" "$OUT"/obj_lib/*.o
if [ -f "$HERE/ref_dump.cpp" ]; then
  g++ -std=c++11 -O2 -w -D_BMI_LIBRARY_ -I"$SRC" "$HERE/ref_dump.cpp" -o "$OUT/ref_dump" \
      -L"$OUT" -lravenref -Wl,-rpath,'$ORIGIN'
fi
if [ -f "$HERE/ref_enkf_dump.cpp" ]; then
  g++ -std=c++11 -O2 -w -D_BMI_LIBRARY_ -I"$SRC" "$HERE/ref_enkf_dump.cpp" -o "$OUT/ref_enkf_dump" \
      -L"$OUT" -lravenref -Wl,-rpath,'$ORIGIN'
fi
echo "build_ref: ok -> $OUT"
