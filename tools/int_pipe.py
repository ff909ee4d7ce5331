import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from conftest import load_package
pkg=load_package(); lib=pkg.load_library(); ctx=pkg.ProverContext(10,lib)
for m,n in [(0,'IMAD'),(1,'IMAD.WIDE'),(2,'Fq mul'),(3,'Fq sqr')]: print(n, round(ctx.bench_int_pipe(m),1),'G/s')
