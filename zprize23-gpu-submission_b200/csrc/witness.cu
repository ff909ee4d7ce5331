// Witness synthesis for the Poseidon-Merkle circuit family ON THE DEVICE.
//
// The reference builds this witness on one CPU thread by running the gadget — `StandardComposer` + HashMap<Variable, F>
// ("Prize 1B/plonk-core/src/constraint_system/hash.rs":20-127 full / partial affine-transform gates,
// "Prize 1B/plonk-hashing/src/poseidon/zprize_constraints.rs":141-265 round structure, "Prize 1B/merkle-tree/src/lib.rs":41-59
// tree walk) — which its README times at 9.4 s per proof, as long as its GPU prover.  The trace is embarrassingly parallel
// inside a tree level: one thread per hash evaluates the 3 key additions and 63 rounds (4 full, 55 partial, 4 full; x^5
// S-box) and writes the 193 gate rows of that hash straight into the prover's resident wire columns, level by level from
// the leaves (14 launches at HEIGHT=15, ~800 dependent Fr products per thread).  Nothing crosses PCIe except the leaves.
//
// Row layout (gate order of the composer, identical to the circuit the prover key was built from):
//   0            zero gate                          (0, 0, 0, 0)
//   1, 2         blinding rows                      (b0..b3), (b4..b7)
//   3            (b4, b5, 0, 0)
//   4 + 193 h …  hash h = nleaves - 2 - node:  3 rows (in_k, 0, in_k + ark_k, 0);  63 x 3 rows (s0, s1, out_j, s2);
//                1 row (s1, node value, 0, 0)       [assert_equal]
//   last         (root, 0, 0, 0)                    public input -root
#include "prover.cuh"

namespace zp {

struct HashParamsDev {
    const fr_t* mds;   // 9: row-major 3 x 3
    const fr_t* ark0;  // 3: keys added before the first round
    const fr_t* rc;    // 63 x 3 round constants
};

__global__ void __launch_bounds__(128) merkle_level_kernel(fr_t* __restrict__ wl, fr_t* __restrict__ wr, fr_t* __restrict__ wo,
                                                           fr_t* __restrict__ w4, fr_t* __restrict__ node, size_t first, size_t count,
                                                           size_t nleaves, HashParamsDev hp) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const size_t i = first + t;  // heap index of the node being hashed
    const fr_t zero = fr_t::zero();
    fr_t in[3] = {zero, load_fr(&node[2 * i + 1]), load_fr(&node[2 * i + 2])};
    size_t row = 4 + 193 * (nleaves - 2 - i);
    fr_t s[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        s[k] = in[k] + load_fr(&hp.ark0[k]);
        store_fr(&wl[row], in[k]);
        store_fr(&wr[row], zero);
        store_fr(&wo[row], s[k]);
        store_fr(&w4[row], zero);
        row++;
    }
    fr_t m[9];
#pragma unroll
    for (int k = 0; k < 9; k++) m[k] = load_fr(&hp.mds[k]);
#pragma unroll 1
    for (int rd = 0; rd < 63; rd++) {
        const bool full = rd < 4 || rd >= 59;
        fr_t p0 = s[0].pow5();
        fr_t p1 = full ? s[1].pow5() : s[1];
        fr_t p2 = full ? s[2].pow5() : s[2];
        fr_t nx[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            // gate value in the composer's order: q_hl a^5 + q_hr|q_r b(^5) + q_h4|q_4 d(^5) + q_c   (hash.rs:23-120)
            nx[j] = m[3 * j] * p0 + m[3 * j + 1] * p1 + m[3 * j + 2] * p2 + load_fr(&hp.rc[3 * rd + j]);
            store_fr(&wl[row], s[0]);
            store_fr(&wr[row], s[1]);
            store_fr(&wo[row], nx[j]);
            store_fr(&w4[row], s[2]);
            row++;
        }
        s[0] = nx[0];
        s[1] = nx[1];
        s[2] = nx[2];
    }
    store_fr(&node[i], s[1]);
    store_fr(&wl[row], s[1]);
    store_fr(&wr[row], s[1]);  // the tree node holds the digest: assert_equal(s1, node)
    store_fr(&wo[row], zero);
    store_fr(&w4[row], zero);
}

__global__ void merkle_frame_kernel(fr_t* wl, fr_t* wr, fr_t* wo, fr_t* w4, const fr_t* blind, const fr_t* node, size_t last_row) {
    if (blockIdx.x || threadIdx.x) return;
    const fr_t zero = fr_t::zero();
    fr_t* w[4] = {wl, wr, wo, w4};
    for (int k = 0; k < 4; k++) {
        store_fr(&w[k][0], zero);
        store_fr(&w[k][1], load_fr(&blind[k]));
        store_fr(&w[k][2], load_fr(&blind[4 + k]));
        store_fr(&w[k][3], k < 2 ? load_fr(&blind[4 + k]) : zero);
        store_fr(&w[k][last_row], k == 0 ? load_fr(&node[0]) : zero);
    }
}

// leaves: 2^(height-1) Fr; params: 9 MDS + 3 + 189 Fr; blinding: 8 Fr (all Montgomery, host).  Leaves the witness resident
// (as upload_witness does) with public input -root at the last gate; root_out (optional) receives the root (Montgomery).
void Prover::synthesize_merkle_witness(int height, const uint64_t* leaves, const uint64_t* params, const uint64_t* blinding,
                                       uint64_t* root_out) {
    if (height < 2 || height > 24) throw std::runtime_error("synthesize_merkle_witness: height out of range");
    const size_t nleaves = (size_t)1 << (height - 1), nnodes = 2 * nleaves - 1;
    const size_t cn = 4 + 193 * (nleaves - 1) + 1;
    if (cn > n) throw std::runtime_error("synthesize_merkle_witness: the circuit does not fit this context's domain");
    ensure_work_buffers(false);
    DevBuf<fr_t> node(nnodes), par(9 + 3 + 189 + 8);
    ZP_CUDA(cudaMemcpyAsync(node.p + (nleaves - 1), leaves, nleaves * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    ZP_CUDA(cudaMemcpyAsync(par.p, params, 201 * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    ZP_CUDA(cudaMemcpyAsync(par.p + 201, blinding, 8 * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    HashParamsDev hp{par.p, par.p + 9, par.p + 12};
    for (int level = height - 2; level >= 0; level--) {
        const size_t first = ((size_t)1 << level) - 1, count = (size_t)1 << level;
        ZP_LAUNCH(merkle_level_kernel, dim3((unsigned)((count + 127) / 128)), dim3(128), 0, st, w_ev[0].p, w_ev[1].p, w_ev[2].p,
                  w_ev[3].p, node.p, first, count, nleaves, hp);
    }
    ZP_LAUNCH(merkle_frame_kernel, dim3(1), dim3(32), 0, st, w_ev[0].p, w_ev[1].p, w_ev[2].p, w_ev[3].p, par.p + 201, node.p, cn - 1);
    for (int k = 0; k < 4; k++)
        if (cn < n) ZP_CUDA(cudaMemsetAsync(w_ev[k].p + cn, 0, (n - cn) * sizeof(fr_t), st));
    ZP_CUDA(cudaMemsetAsync(qlk_ev.p, 0, n * sizeof(fr_t), st));
    fr_t root;
    ZP_CUDA(cudaMemcpyAsync(&root, node.p, sizeof(fr_t), cudaMemcpyDeviceToHost, st));
    ZP_CUDA(cudaStreamSynchronize(st));
    if (root_out) memcpy(root_out, root.l, 32);
    host::Fr neg = host::to_host(root).neg();
    neg.to_canonical(wit_pi);  // CircuitC.pi convention: canonical form
    wit_pi_pos = cn - 1;
    wit_n = cn;
    wit_lookup_on = !table_zero;  // q_lookup == 0: lookups stay on only when the key carries a table
    if (wit_lookup_on) ensure_work_buffers(true);
}

void Prover::read_witness(int k, uint64_t* out) {
    if (k < 0 || k > 3 || !wit_n) throw std::runtime_error("read_witness: no witness resident");
    ZP_CUDA(cudaMemcpy(out, w_ev[k].p, wit_n * sizeof(fr_t), cudaMemcpyDeviceToHost));
}

}  // namespace zp
