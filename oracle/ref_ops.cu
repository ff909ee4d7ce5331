// ORACLE — TEST / MEASUREMENT INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin extern "C" driver over the REFERENCE's own operator API, compiled against the reference's headers where they
// lie ("Prize 1B/plonk-core/lib/PLONK/utils/function.cuh":88-129) and linked with oracle/_ref/libzprize_ref.so (the
// reference's native code built unmodified for sm_100 by build_pnp_ref.sh).  It lets tools/bench_vs_pnp_reference.py
// time `Ntt/Intt/Ntt_coset/Intt_coset::forward` (function.cu:249-273, sppark-derived NTT) and `multi_scalar_mult`
// (function.cu:275-290, sppark-derived Pippenger + CPU collect) on the same B200 next to our kernels, and check that
// both produce the same bytes.  Inputs are made device-resident before the timed loop, like our `zp_bench_*` entry
// points; everything the reference does inside one call (allocation, parameter blobs, D2H + CPU tail of the MSM) is
// part of its operator and is timed.
#include "PLONK/utils/function.cuh"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

namespace {
double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// the reference's wrappers allocate / free device memory inside every call, so single calls fluctuate: every call is timed
// on its own (device synchronised on both sides) and the minimum and the median are reported
template <class F>
void time_calls(int iters, F&& call, double* ms_min, double* ms_median) {
    std::vector<double> t;
    for (int i = 0; i < iters; i++) {
        cudaDeviceSynchronize();
        double t0 = now_ms();
        call();
        cudaDeviceSynchronize();
        t.push_back(now_ms() - t0);
    }
    std::sort(t.begin(), t.end());
    *ms_min = t.front();
    *ms_median = t[t.size() / 2];
}
}  // namespace

// kind: 0 NTT, 1 iNTT, 2 coset NTT of n coefficients zero-padded to 8n (out: 8n), 3 coset iNTT (in/out: n)
extern "C" int ref_ops_ntt(int kind, int logn, const uint64_t* in, uint64_t* out, int iters, double* ms, double* ms_median) {
    const size_t n = (size_t)1 << logn;
    SyncedMemory x(n * 4 * sizeof(uint64_t));
    memcpy(x.mutable_cpu_data(), in, n * 32);
    x.mutable_gpu_data();  // host -> device before the timed region
    cudaDeviceSynchronize();
    SyncedMemory y((size_t)0);
    // the reference builds its transform objects with fr::TWO_ADICITY = 32 (gen_proof.cuh:30, quotient.cu:170,374)
    if (kind == 0) {
        Ntt t(32);
        y = t.forward(x);
        time_calls(iters, [&] { y = t.forward(x); }, ms, ms_median);
    } else if (kind == 1) {
        Intt t(32);
        y = t.forward(x);
        time_calls(iters, [&] { y = t.forward(x); }, ms, ms_median);
    } else if (kind == 2) {
        Ntt_coset t(32, (int)(8 * n));
        y = t.forward(x);
        time_calls(iters, [&] { y = t.forward(x); }, ms, ms_median);
    } else {
        Intt_coset t(32);
        y = t.forward(x);
        time_calls(iters, [&] { y = t.forward(x); }, ms, ms_median);
    }
    if (out) memcpy(out, y.mutable_cpu_data(), y.size());
    return (int)(y.size() / 32 >> 0 ? 0 : 1);
}

// points: n affine points (x || y, Montgomery Fq); scalars: n Montgomery Fr (converted with the reference's own to_base,
// as its KZG commit does: kzg10.cu:31-33); out: Jacobian X, Y, Z (18 u64, Montgomery) as multi_scalar_mult returns it
extern "C" int ref_ops_msm(size_t n, const uint64_t* points, const uint64_t* scalars, uint64_t* out, int iters, double* ms,
                           double* ms_median) {
    SyncedMemory p(n * 12 * sizeof(uint64_t)), s(n * 4 * sizeof(uint64_t));
    memcpy(p.mutable_cpu_data(), points, n * 96);
    memcpy(s.mutable_cpu_data(), scalars, n * 32);
    p.mutable_gpu_data();
    s.mutable_gpu_data();
    SyncedMemory sb = to_base(s);
    sb.mutable_gpu_data();
    cudaDeviceSynchronize();
    SyncedMemory r = multi_scalar_mult(p, sb);
    cudaDeviceSynchronize();
    time_calls(iters, [&] { r = multi_scalar_mult(p, sb); }, ms, ms_median);
    memcpy(out, r.mutable_cpu_data(), 18 * sizeof(uint64_t));
    return 0;
}
