"""Pipe microbenchmarks on the GPU (zp_bench_int_pipe): the roofline denominators of DESIGN.md and the round-2 multiplier study."""
import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from conftest import load_package
pkg = load_package(); lib = pkg.load_library(); ctx = pkg.ProverContext(10, lib)
for m, n in [(0, 'IMAD (mad.lo.u32)'), (1, 'IMAD.WIDE'), (2, 'Fq Montgomery mul'), (3, 'Fq Montgomery sqr'), (4, 'FP64 FMA'),
             (5, 'FP64 FMA with 1:1 IMAD beside it (FP64 ops counted)'), (6, 'IADD3 (ALU pipe)')]:
    print(n, round(ctx.bench_int_pipe(m), 1), 'G/s', flush=True)
