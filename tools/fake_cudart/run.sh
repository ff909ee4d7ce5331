#!/usr/bin/env bash
# Builds the fake CUDA runtime and runs the reference's native prover (unmodified and patched) on it at HEIGHT=5.
# No GPU needed.  Usage: tools/fake_cudart/run.sh [height]
set -uo pipefail
cd "$(dirname "$0")/../.."
H="${1:-5}"
D=/tmp/zp_fake_cudart
mkdir -p "$D"
g++ -O1 -g -fPIC -shared -I/usr/local/cuda/include tools/fake_cudart/fake_cudart.cpp \
    -Wl,--version-script=tools/fake_cudart/ver.map -Wl,-soname,libcudart.so.12 -o "$D/libcudart.so.12" || exit 1
for lib in libzprize_ref.so libzprize_ref_patched.so; do
  for node in 0 56; do
    echo "== $lib HEIGHT=$H SHIM_NODE=$node"
    if [ "$node" = 0 ]; then unset SHIM_NODE; else export SHIM_NODE=$node; fi
    LD_LIBRARY_PATH="$D" ZP_NO_SEGV_TRACE=1 timeout 900 python tools/run_pnp_reference.py --height "$H" --out "$D/proof.npy" --lib "$lib" > "$D/run.log" 2>&1
    echo "exit code $?"
    grep "shim\]\|reference gen_proof" "$D/run.log" | sort | uniq -c
    grep -m1 -A12 "shim\] PINNED" "$D/run.log" | grep "libzprize_ref" | head -6
  done
done
