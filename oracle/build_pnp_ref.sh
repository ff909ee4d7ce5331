#!/usr/bin/env bash
# ORACLE — TEST INFRASTRUCTURE ONLY.
# Compiles the reference's OWN native prover (PNP's C++/CUDA `lib/`, unmodified, taken where it lies under
# /root/reference — never copied) for sm_100 into oracle/_ref/libzprize_ref.so, following the recipe of
# "Prize 1B/plonk-core/build.rs":36-103 (blst: gcc -O2 -mno-avx; everything else: nvcc -std=c++17 -O3, here
# -arch=sm_100 instead of sm_80).  The result exports the reference's `gen_proof` symbol and is used ONLY by
# tests/test_gpu_vs_pnp_reference.py to cross-run the reference on the GPU box against the oracle and against
# our library (SURVEY §8c pin (5); valid for Merkle-shaped inputs).
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
  if [ "$(basename "$f")" = "zk_function.cu" ]; then
    # DOCUMENTED PATCH (second library only; the first one stays unmodified): the reference's split_tx_poly
    # (lib/PLONK/utils/zk_function.cu:38-48) copies 8 slices of n elements out of t_poly without checking t_poly's size; on
    # sm_100 that cudaMemcpy faults above HEIGHT=4 (profiles/r02b_pnp_reference_crash_backtrace.log).  The patched copy
    # (transient, under _ref/pnp_obj, deleted below) clamps every slice to what t_poly holds, zero-fills the rest and
    # reports the sizes, so that the reference can be timed at HEIGHT=15 on the same B200.
    d="$(dirname "$f")"
    python3 - "$f" _ref/pnp_obj/zk_function_patched.cu <<'PYEOF'
import sys
src = open(sys.argv[1]).read()
old = "caffe_gpu_memcpy(t_.size(), t_gpu + i*t_.size(), t_x_gpu);"
new = ("{ size_t off = (size_t)i * t_.size(); size_t avail = off < t_poly.size() ? t_poly.size() - off : 0; "
       "size_t cnt = avail < t_.size() ? avail : t_.size(); "
       "cudaPointerAttributes pa, pb; cudaPointerGetAttributes(&pa, (char*)t_gpu + off); cudaPointerGetAttributes(&pb, t_x_gpu); "
       "if (i == 0) fprintf(stderr, \"[ref-patch] split_tx_poly: t_poly holds %zu bytes, 8 slices need %zu\\n\", t_poly.size(), 8 * t_.size()); "
       "fprintf(stderr, \"[ref-patch] slice %d: src %p (memory type %d) dst %p (memory type %d) bytes %zu\\n\", i, (char*)t_gpu + off, (int)pa.type, t_x_gpu, (int)pb.type, cnt); "
       "cudaError_t e1 = cudaMemset(t_x_gpu, 0, t_.size()); "
       "cudaError_t e2 = cnt ? cudaMemcpy(t_x_gpu, (char*)t_gpu + off, cnt, cudaMemcpyDeviceToDevice) : cudaSuccess; "
       "if (e1 != cudaSuccess || e2 != cudaSuccess) fprintf(stderr, \"[ref-patch] memset: %s, memcpy: %s\\n\", cudaGetErrorString(e1), cudaGetErrorString(e2)); }")
assert old in src
open(sys.argv[2], "w").write(src.replace(old, new))
PYEOF
    ( "$NVCC" -std=c++17 -O3 -arch=sm_100 -Xcompiler -fPIC -ccbin g++ -w -include cstdint -I"$REF/blst/include" -I"$d" \
        -c _ref/pnp_obj/zk_function_patched.cu -o _ref/pnp_obj/patched_zk_function.o ) &
    pids+=($!)
    echo "$o" > _ref/pnp_obj/zk_function_obj_name
  fi
  if (( ${#pids[@]} >= ${ZP_JOBS:-8} )); then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
done < <(find "$REF/PLONK" "$REF/caffe" "$REF/hello.cu" \( -name '*.cu' -o -name '*.cpp' \) -print0)
wait || true
ORIG_ZK="$(cat _ref/pnp_obj/zk_function_obj_name)"
mv _ref/pnp_obj/patched_zk_function.o _ref/pnp_obj/patched_zk_function.obj
"$NVCC" -shared -arch=sm_100 -o "$OUT" _ref/pnp_obj/*.o -lcudart -lpthread
mv "$ORIG_ZK" "$ORIG_ZK.orig"
mv _ref/pnp_obj/patched_zk_function.obj _ref/pnp_obj/patched_zk_function.o
"$NVCC" -shared -arch=sm_100 -o "$PATCHED" _ref/pnp_obj/*.o -lcudart -lpthread
rm -rf _ref/pnp_obj
echo "built $OUT"
build_ops
