#!/usr/bin/env bash
# Builds the product sources against the CPU emulation layer with AddressSanitizer and runs tools/asan_emu/run.py.  No GPU.
set -euo pipefail
cd "$(dirname "$0")/../.."
R="$PWD"; C="$R/zprize23-gpu-submission_b200/csrc"; E="$R/tests/emu"; D=/tmp/zp_asan_emu
mkdir -p "$D"
for s in ntt msm poly wiring witness prover verifier capi; do
  g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -DZP_EMU -include "$E/cuda_emu.h" -Wno-unused-function \
      -Wno-attributes -Wno-unknown-pragmas -I"$R/include" -x c++ -c "$C/$s.cu" -o "$D/$s.o" 2>/dev/null &
done
g++ -std=c++17 -O1 -g -fsanitize=address -fPIC -c "$E/cuda_emu.cpp" -o "$D/cuda_emu.o" &
wait
g++ -shared -fsanitize=address -o "$D/libzprize_emu_asan.so" "$D"/*.o
for rounds in 1 2 4; do
  LD_PRELOAD="$(gcc -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    python tools/asan_emu/run.py "$D/libzprize_emu_asan.so" "$rounds" 2>&1 | grep -v "doesn't fully support makecontext"
done
echo "AddressSanitizer: no report"
