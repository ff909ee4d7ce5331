#!/usr/bin/env bash
# ORACLE — TEST INFRASTRUCTURE ONLY.
# Compiles the reference's OWN native prover (PNP's C++/CUDA `lib/`, unmodified, taken where it lies under
# /root/reference — never copied) for sm_100 into oracle/_ref/libzprize_ref.so, following the recipe of
# "Prize 1B/plonk-core/build.rs":36-103 (blst: gcc -O2 -mno-avx; everything else: nvcc -std=c++17 -O3, here
# -arch=sm_100 instead of sm_80).  The result exports the reference's `gen_proof` symbol and is used ONLY by
# tests/test_gpu_vs_pnp_reference.py to cross-run the reference on the GPU box against the oracle and against
# our library (SURVEY §8c pin (5); valid for Merkle-shaped inputs).  A second library,
# oracle/_ref/libzprize_ref_patched.so, is the same build with three translation units patched at build time (below): the
# unmodified one dies above HEIGHT=4 from its own double destruction of shared buffers; the patched one proves HEIGHT=15.
set -euo pipefail
cd "$(dirname "$0")"
REF="${ZP_REFERENCE_ROOT:-/root/reference}/Prize 1B/plonk-core/lib"
[ -d "$REF" ] || { echo "reference not present; skipping"; exit 0; }
mkdir -p _ref/pnp_obj
OUT=_ref/libzprize_ref.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# extern "C" driver over the reference's operator API (ref_ops.cu), linked against the library above
build_ops() {
  if [ ! -f _ref/libref_ops.so ] || [ ref_ops.cu -nt _ref/libref_ops.so ]; then
    "$NVCC" -std=c++17 -O3 -arch=sm_100 -Xcompiler -fPIC -ccbin g++ -w -include cstdint -I"$REF" -I"$REF/blst/include" \
        -shared ref_ops.cu -o _ref/libref_ops.so -L_ref -lzprize_ref -Xlinker -rpath -Xlinker '$ORIGIN' -lcudart
    echo "built _ref/libref_ops.so"
  fi
}
PATCHED=_ref/libzprize_ref_patched.so
[ -f "$OUT" ] && [ -f "$PATCHED" ] && { echo "$OUT exists"; build_ops; exit 0; }
rm -rf _ref/pnp_obj && mkdir -p _ref/pnp_obj   # leftovers of an interrupted build
gcc -O2 -mno-avx -fno-builtin -Wno-unused-function -fPIC -D__BLST_PORTABLE__ -I"$REF/blst/include" \
    -c "$REF/blst/src/server.c" -o _ref/pnp_obj/blst_server.o
gcc -O2 -fPIC -c "$REF/blst/src/assembly.S" -o _ref/pnp_obj/blst_asm.o
i=0
pids=()
while IFS= read -r -d '' f; do
  i=$((i+1))
  o="_ref/pnp_obj/tu_$i.o"
  ( "$NVCC" -std=c++17 -O3 -arch=sm_100 -Xcompiler -fPIC -ccbin g++ -w -include cstdint -I"$REF/blst/include" -c "$f" -o "$o" ) &
  pids+=($!)
  case "$(basename "$f")" in quotient.cu|point.cu|msmcollect.cpp)
    # DOCUMENTED PATCH (second library only; the first one stays unmodified).  The reference calls destructors explicitly on
    # locals that are destroyed AGAIN at scope exit — `arithmetic_evals.~Arithmetic(); selector_evals.~Selectors();
    # permutation_evals.~Permutation();` (lib/PLONK/src/plonk_core/src/proof_system/quotient.cu:277-278,323) and
    # `midN.~SyncedMemory();` (lib/PLONK/src/point.cu:128-159,224-253).  Their members are std::shared_ptr (caffe/interface.hpp:9),
    # so every buffer is released twice: undefined behaviour that ends, on this box, with the quotient polynomial's device
    # buffer freed before split_tx_poly reads it (SIGSEGV above HEIGHT=4, profiles/r02b_pnp_reference_crash_backtrace.log;
    # at HEIGHT=4 the stale 512 KiB block is still mapped).  The patched copy (transient, under _ref/pnp_obj, deleted below)
    # drops the three struct destructor calls and turns `x.~SyncedMemory();` into the well-defined `x = SyncedMemory();`
    # — the same early release, once.  Second fix, same library: msm_collect_cpu (lib/PLONK/utils/zkp/cpu/msmcollect.cpp:35)
    # allocates its 144-byte Jacobian result as `SyncedMemory out(3 * fq_LIMBS, false)` = 18 BYTES of pinned host memory and
    # writes 126 bytes past its end on every MSM (seen with tools/fake_cudart: "18-byte buffer written 126 bytes past its
    # end", 33 times per proof); the patched copy sizes it 3 * fq_LIMBS * sizeof(uint64_t).  Nothing else changes, so the
    # reference can be run and timed at HEIGHT=15.
    d="$(dirname "$f")"
    b="$(basename "$f")"; x="${b##*.}"; b="${b%.*}"
    python3 - "$f" "_ref/pnp_obj/${b}_patched.$x" <<'PYEOF'
import re, sys
src = open(sys.argv[1]).read()
out, n1 = re.subn(r"^\s*\w+\.~(Arithmetic|Selectors|Permutation)\(\);\s*$", "", src, flags=re.M)
out, n2 = re.subn(r"(\w+)\.~SyncedMemory\(\);", r"\1 = SyncedMemory();", out)
out, n3 = re.subn(r"SyncedMemory out\(3 \* fq_LIMBS, false\);", "SyncedMemory out(3 * fq_LIMBS * sizeof(uint64_t), false);", out)
assert n1 + n2 + n3 > 0, "nothing to patch in " + sys.argv[1]
sys.stderr.write("[build_pnp_ref] %s: %d struct destructor calls dropped, %d SyncedMemory destructor calls turned into resets, %d result buffers resized\n" % (sys.argv[1].split("/")[-1], n1, n2, n3))
open(sys.argv[2], "w").write(out)
PYEOF
    ( "$NVCC" -std=c++17 -O3 -arch=sm_100 -Xcompiler -fPIC -ccbin g++ -w -include cstdint -I"$REF/blst/include" -I"$d" \
        -c "_ref/pnp_obj/${b}_patched.$x" -o "_ref/pnp_obj/patched_${b}.obj" ) &
    pids+=($!)
    echo "$o" >> _ref/pnp_obj/patched_originals
    ;;
  esac
  if (( ${#pids[@]} >= ${ZP_JOBS:-8} )); then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
done < <(find "$REF/PLONK" "$REF/caffe" "$REF/hello.cu" \( -name '*.cu' -o -name '*.cpp' \) -print0)
wait || true
"$NVCC" -shared -arch=sm_100 -o "$OUT" _ref/pnp_obj/*.o -lcudart -lpthread
# second library: the same objects with the patched translation units in place of their originals
while IFS= read -r orig; do mv "$orig" "$orig.orig"; done < _ref/pnp_obj/patched_originals
for p in _ref/pnp_obj/patched_*.obj; do mv "$p" "${p%.obj}.o"; done
"$NVCC" -shared -arch=sm_100 -o "$PATCHED" _ref/pnp_obj/*.o -lcudart -lpthread
rm -rf _ref/pnp_obj
echo "built $OUT"
build_ops
