#!/usr/bin/env bash
# ORACLE — TEST INFRASTRUCTURE ONLY.
# Builds (1) oracle/liboracle.so, our CPU restatement, and (2) when /root/reference is present,
# oracle/_ref/: the reference's OWN sources compiled where they lie (never copied into this repo):
#   libref_blst.so   <- "Prize 1B/plonk-core/lib/blst/src/server.c" + "assembly.S"
#                       (recipe of "Prize 1B/plonk-core/build.rs":36-54: -O2 -mno-avx -fno-builtin)
#   libref_strobe.so <- "Prize 1B/plonk-core/lib/PLONK/src/transcript/strobe.cpp" + oracle/ref_shim.cpp
# oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot.
set -euo pipefail
cd "$(dirname "$0")"
CXX=${ZP_CXX:-g++}
CC=${ZP_CC:-gcc}
if [ ! -f liboracle.so ] || [ -n "$(find . -maxdepth 1 \( -name '*.hpp' -o -name '*.cpp' \) -newer liboracle.so)" ]; then
  $CXX -std=c++17 -O2 -fopenmp -fPIC -shared -Wall -Wno-unused-function -o liboracle.so oracle_capi.cpp -ldl
fi
REF="${ZP_REFERENCE_ROOT:-/root/reference}/Prize 1B/plonk-core/lib"
if [ -d "$REF" ]; then
  mkdir -p _ref
  if [ ! -f _ref/libref_blst.so ]; then
    $CC -O2 -mno-avx -fno-builtin -Wno-unused-function -fPIC -shared -D__BLST_PORTABLE__ \
        -I"$REF/blst/include" "$REF/blst/src/server.c" "$REF/blst/src/assembly.S" -o _ref/libref_blst.so
  fi
  if [ ! -f _ref/libref_strobe.so ] || [ ref_shim.cpp -nt _ref/libref_strobe.so ]; then
    $CXX -std=c++17 -O2 -fPIC -shared -I"$REF/PLONK/src/transcript" \
        "$REF/PLONK/src/transcript/strobe.cpp" ref_shim.cpp -o _ref/libref_strobe.so
  fi
fi
if [ ! -f libsegv_trace.so ] || [ segv_trace.c -nt libsegv_trace.so ]; then
  $CC -O1 -g -fPIC -shared -rdynamic segv_trace.c -o libsegv_trace.so
fi
echo "oracle build ok"
