/* C-ABI of the B200-native PLONK prover hot path (libzprize_b200.so).
 *
 * Drop-in boundary: `gen_proof` below has exactly the signature the reference binds from Rust
 *   extern "C" { pub fn gen_proof(circuit: CircuitC, pk: ProverKeyC, ck: CommitKeyC) -> ProofC; }
 *   ("Prize 1B/plonk-core/src/lib.rs":237-239; C++ side "Prize 1B/plonk-core/lib/hello.cu":4-7), and the
 * structs are field-for-field the `#[repr(C)]` structs of lib.rs:53-235 (mirrored in C by the reference at
 * "Prize 1B/plonk-core/lib/PLONK/src/structure.cuh":7-329).  Everything else in this header is an
 * extension (resident prover context, per-operator entry points for the NTT / MSM sweeps) — plain
 * pointers and sizes only.
 *
 * Data conventions (all identical to the reference FFI):
 *   Fr  = 4 x u64 little-endian limbs, Montgomery form (R = 2^256), canonical (< r)
 *   Fq  = 6 x u64 little-endian limbs, Montgomery form (R = 2^384)
 *   G1 affine = x || y (12 x u64), no infinity flag; infinity in outputs is (x = 0, y = Mont(1))
 *   CircuitC.pi is ONE Fr in canonical (non-Montgomery) form (prover.rs:721-725)
 */
#ifndef ZPRIZE_B200_H
#define ZPRIZE_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- lib.rs:53-118 ------------------------------------------------------------------------ */
typedef struct { uint64_t a_eval[4], b_eval[4], c_eval[4], d_eval[4]; } WireEvaluationsC;
typedef struct { uint64_t left_sigma_eval[4], right_sigma_eval[4], out_sigma_eval[4], permutation_eval[4]; } PermutationEvaluationsC;
typedef struct {
    uint64_t q_lookup_eval[4], z2_next_eval[4], h1_eval[4], h1_next_eval[4], h2_eval[4], f_eval[4], table_eval[4],
        table_next_eval[4];
} LookupEvaluationsC;
typedef struct {
    uint64_t q_arith_eval[4], q_c_eval[4], q_l_eval[4], q_r_eval[4], q_hl_eval[4], q_hr_eval[4], q_h4_eval[4], a_next_eval[4],
        b_next_eval[4], d_next_eval[4];
} CustomEvaluationsC;
typedef struct {
    WireEvaluationsC wire_evals;
    PermutationEvaluationsC perm_evals;
    LookupEvaluationsC lookup_evals;
    CustomEvaluationsC custom_evals;
} ProofEvaluationsC;

/* lib.rs:231-235 */
typedef struct { uint64_t x[6]; uint64_t y[6]; } CommitmentC;

/* lib.rs:120-142 */
typedef struct {
    CommitmentC a_comm, b_comm, c_comm, d_comm, z_comm, f_comm, h_1_comm, h_2_comm, z_2_comm;
    CommitmentC t_1_comm, t_2_comm, t_3_comm, t_4_comm, t_5_comm, t_6_comm, t_7_comm, t_8_comm;
    CommitmentC aw_opening, saw_opening;
    ProofEvaluationsC evaluations;
} ProofC;

/* lib.rs:144-155 */
typedef struct {
    uint64_t n;               /* unpadded gate count cs.n */
    uint64_t lookup_len;      /* rows of the lookup table */
    uint64_t intended_pi_pos; /* position of the single public input */
    uint64_t* q_lookup;       /* n Fr */
    uint64_t* pi;             /* 1 Fr, canonical form */
    uint64_t* w_l;            /* n Fr each */
    uint64_t* w_r;
    uint64_t* w_o;
    uint64_t* w_4;
} CircuitC;

/* lib.rs:157-223 — *_coeffs: <= N Fr (arkworks trims trailing zeros; see zp_prover_load_pk for how
 * identically-zero polynomials are handled), *_evals: 8N Fr on the coset g*H_8N, natural order. */
typedef struct {
    uint64_t *q_m_coeffs, *q_m_evals;
    uint64_t *q_l_coeffs, *q_l_evals;
    uint64_t *q_r_coeffs, *q_r_evals;
    uint64_t *q_o_coeffs, *q_o_evals;
    uint64_t *q_4_coeffs, *q_4_evals;
    uint64_t *q_c_coeffs, *q_c_evals;
    uint64_t *q_hl_coeffs, *q_hl_evals;
    uint64_t *q_hr_coeffs, *q_hr_evals;
    uint64_t *q_h4_coeffs, *q_h4_evals;
    uint64_t *q_arith_coeffs, *q_arith_evals;
    uint64_t *range_selector_coeffs, *range_selector_evals;
    uint64_t *logic_selector_coeffs, *logic_selector_evals;
    uint64_t *fixed_group_add_selector_coeffs, *fixed_group_add_selector_evals;
    uint64_t *variable_group_add_selector_coeffs, *variable_group_add_selector_evals;
    uint64_t *q_lookup_coeffs, *q_lookup_evals;
    uint64_t *table1, *table2, *table3, *table4; /* N Fr each */
    uint64_t *left_sigma_coeffs, *left_sigma_evals;
    uint64_t *right_sigma_coeffs, *right_sigma_evals;
    uint64_t *out_sigma_coeffs, *out_sigma_evals;
    uint64_t *fourth_sigma_coeffs, *fourth_sigma_evals;
    uint64_t* linear_evaluations; /* 8N — closed form, never read */
    uint64_t* v_h_coset_8n;       /* 8N — closed form, never read */
} ProverKeyC;

/* lib.rs:225-229 */
typedef struct {
    const uint64_t* powers_of_g;       /* >= N affine points */
    const uint64_t* powers_of_gamma_g; /* unused (hiding is off) */
} CommitKeyC;

/* ---- the drop-in symbol (lib.rs:237-239) ------------------------------------------------------
 * Uploads the prover key and SRS on the first call and keeps them resident in HBM for later calls with the SAME KEY
 * CONTENT (domain size + fingerprint of every pk array, the tables and the SRS; addresses are irrelevant, so the
 * reference's `pk.clone()`-per-proof pattern, benches/pnp_bench.rs:70, hits the cache).  ZPRIZE_B200_PK_CACHE = "0"
 * disables caching, "full" fingerprints every word instead of a strided sample; see INTEGRATION.md.
 * Like the reference it has no error channel: a failure (CUDA error, domain outside [2^6, 2^23], public-input position
 * outside the domain) prints a message and exits(1) (lib/caffe/common.hpp:23-30). */
ProofC gen_proof(CircuitC circuit, ProverKeyC pk, CommitKeyC ck);
/* Drops the context gen_proof keeps resident (frees its HBM); the next call uploads again. */
void zp_gen_proof_invalidate(void);

/* ---- resident prover context (extension) ------------------------------------------------------ */
typedef struct zp_prover zp_prover;

/* last error message of the calling thread ("" if none) */
const char* zp_last_error(void);
/* number of CUDA kernels launched by the library so far (process-wide) */
uint64_t zp_launch_count(void);
/* 1 if a CUDA device is usable */
int zp_device_available(void);

/* Context for domain size N = 2^log_n on the current CUDA device. NULL on error.  log_n in [6, 23] for proving (the 8N
 * extended domain must fit 2^26); log_n in (23, 26] gives an operator-only context (SRS, MSM and NTT entry points). */
zp_prover* zp_prover_create(int log_n);
void zp_prover_destroy(zp_prover* p);
/* Run all work of this context on the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream)
 * instead of the context's own stream, so that the caller's events bracket the prover's kernels. */
int zp_prover_set_stream(zp_prover* p, void* cuda_stream);
/* cudaProfilerStart (1) / cudaProfilerStop (0): lets `ncu --profile-from-start off` capture one proof. */
int zp_profiler_range(int start);
/* transcript label; default "Merkle tree" (gen_proof.cuh:19-20) */
int zp_prover_set_label(zp_prover* p, const char* label);
/* SRS: n_points affine points (CommitKeyC.powers_of_g layout), n_points >= N. */
int zp_prover_load_srs(zp_prover* p, const uint64_t* powers_of_g, size_t n_points);
/* Insecure test SRS generated on the device: powers_of_g[i] = tau^i * G (tau: Montgomery Fr). */
int zp_prover_generate_srs(zp_prover* p, const uint64_t* tau, size_t n_points);
/* Copy the resident SRS back (n_points * 12 u64). */
int zp_prover_read_srs(zp_prover* p, uint64_t* out, size_t n_points);
/* Upload a prover key given in the reference FFI layout.  coeff_len[19] gives the readable length of
 * each *_coeffs array in ProverKeyC order (q_m .. q_lookup, then the four sigmas); pass NULL to apply
 * the reference's convention (gen_proof.cuh:61-62,277-278,319-329): q_m, range, logic, fixed, variable
 * and q_lookup coefficient arrays are NOT read (they are recovered from the *_evals arrays), all
 * others hold exactly N elements. */
int zp_prover_load_pk(zp_prover* p, const ProverKeyC* pk, const uint64_t* coeff_len);
/* Build the prover key ON THE DEVICE from the circuit description (preprocessing,
 * "Prize 1B/plonk-core/src/proof_system/preprocess.rs":162-295): selector_evals[19] are the 15 selector
 * columns followed by the 4 sigma columns as evaluations on H (N Fr each, host memory, NULL = all
 * zero); tables[4] the padded lookup columns (N Fr each, NULL = all zero). */
int zp_prover_preprocess(zp_prover* p, const uint64_t* const* selector_evals, const uint64_t* const* tables);
/* The same with the PERMUTATION part of preprocessing on the device too (permutation/mod.rs:101-215): instead of the four
 * sigma columns the caller passes the circuit's wire map — the reference's `Permutation::variable_map` flattened in
 * insertion order: m entries, vars[i] = variable id (< n_vars), cells[i] = (gate << 2) | wire with wire 0 = left, 1 = right,
 * 2 = output, 3 = fourth.  Every cell maps to the next cell of its variable (last -> first), unmapped cells to themselves;
 * sigma_k(omega^i) = K_wire' * omega^gate' (K = 1, 7, 13, 17).  selector_evals15: the 15 selector columns (NULL = zero). */
int zp_prover_preprocess_wiring(zp_prover* p, const uint64_t* const* selector_evals15, const uint32_t* vars, const uint32_t* cells,
                                size_t m, uint32_t n_vars, const uint64_t* const* tables);
/* Operator form: only the four sigma columns (N Fr each, host) from the wire map. */
int zp_sigma_from_wiring_host(zp_prover* p, const uint32_t* vars, const uint32_t* cells, size_t m, uint32_t n_vars,
                              uint64_t* const* sigma_out);
/* Copy one prover-key polynomial back in the FFI layout: index 0..18 in ProverKeyC order (coeffs_out: N Fr, evals_out:
 * 8N Fr, either may be NULL), index 19..22 = lookup table columns (N Fr through coeffs_out).  With zp_prover_preprocess
 * this is the device twin of `preprocess_prover` (preprocess.rs:162-295) producing the ProverKeyC arrays. */
int zp_prover_read_pk(zp_prover* p, int index, uint64_t* coeffs_out, uint64_t* evals_out);
/* Commitments to the 19 prover-key polynomials + 4 table polynomials (verifier key), 23 * 12 u64. */
int zp_prover_verifier_key(zp_prover* p, uint64_t* out_commitments);
/* One proof with the resident key.  Host pointers in `circuit`; returns 0 on success. */
int zp_prover_prove(zp_prover* p, const CircuitC* circuit, ProofC* out);
/* Split form of zp_prover_prove for kernel-only timing: upload the witness once (host -> HBM), then prove
 * any number of times with every input already resident. */
int zp_prover_upload_witness(zp_prover* p, const CircuitC* circuit);
int zp_prover_prove_resident(zp_prover* p, ProofC* out);
/* Witness synthesis ON THE DEVICE for the Poseidon-Merkle circuit family (the reference runs the gadget on one CPU thread,
 * 9.4 s at HEIGHT=15: constraint_system/hash.rs:20-127, plonk-hashing/src/poseidon/zprize_constraints.rs:141-265,
 * merkle-tree/src/lib.rs:41-59): leaves = 2^(height-1) Fr, hash_params = 3x3 MDS (row-major) || 3 pre-round keys || 63x3 round
 * constants, blinding = the 8 values of the two blinding rows (all Montgomery).  Leaves the four wire columns resident
 * (cs.n = 4 + 193 (2^(height-1) - 1) + 1 rows, public input -root at the last gate) for zp_prover_prove_resident;
 * root_out (optional) receives the root. */
int zp_prover_synthesize_merkle_witness(zp_prover* p, int height, const uint64_t* leaves, const uint64_t* hash_params,
                                        const uint64_t* blinding, uint64_t* root_out);
/* rows of the resident witness / copy of wire column 0..3 (rows x 4 u64) back to the host */
uint64_t zp_prover_witness_rows(zp_prover* p);
int zp_prover_read_witness(zp_prover* p, int wire, uint64_t* out);
/* Per-proof statistics of the MSM bucket-accumulation kernel (the dominant kernel): enable, then after a
 * proof read out4 = { sum of kernel ms, launches, algorithmic 32-bit multiply-adds (10*588*M*W),
 * sum of all MSM kernel ms }. */
int zp_prover_collect_msm_stats(zp_prover* p, int enable);
/* out9: bucket-accumulation ms (batch-affine rounds + XYZZ accumulate), MSM pipelines launched, algorithmic mads
 * (10*588*M*W), all MSM stages ms, mads actually issued (estimate), commitments produced, and for the dominant kernel
 * ba_down0_kernel: device ms, affine additions performed, launches */
int zp_prover_msm_stats(zp_prover* p, double* out9);
/* Multi-GPU sharding of the commitments (one process per GPU, every rank holds the same key and witness):
 * rank r computes each MSM over points [r*ceil(n/world), (r+1)*ceil(n/world)) only and the partial sums
 * are exchanged with `allgather(user, send, recv, bytes_per_rank)` — recv holds world * bytes_per_rank
 * bytes in rank order; return 0 on success.  world = 1 disables sharding. */
typedef int (*zp_allgather_fn)(void* user, const void* send, void* recv, size_t bytes_per_rank);
int zp_prover_set_shard(zp_prover* p, int rank, int world, zp_allgather_fn allgather, void* user);
/* Optional second hook for the sharded prover: a broadcast of DEVICE memory (`bytes` at `dev_ptr`, from rank `root`
 * to all ranks, ordered after prior work on the prover's stream — e.g. torch.distributed.broadcast / ncclBroadcast).
 * When set, the independent 8N coset NTTs of round 4 are computed by different ranks and broadcast over NVLink, and
 * the fused quotient pass is split by index range (each rank's slice broadcast to the others). */
typedef int (*zp_dev_broadcast_fn)(void* user, void* dev_ptr, size_t bytes, int root);
int zp_prover_set_device_broadcast(zp_prover* p, zp_dev_broadcast_fn bcast, void* user);
/* Optional third hook: IN-PLACE all-gather of device memory — rank r's block [r * bytes_per_rank, (r + 1) * bytes_per_rank) of
 * `dev_base` is sent to every rank, ordered on the prover's stream (ncclAllGather / torch all_gather_into_tensor).  When
 * set it replaces the per-owner broadcasts of the witness slices and of the per-coset quotient coefficients by ONE
 * collective each. */
typedef int (*zp_dev_allgather_fn)(void* user, void* dev_base, size_t bytes_per_rank);
int zp_prover_set_device_allgather(zp_prover* p, zp_dev_allgather_fn allgather, void* user);
/* Device-milliseconds of the phases of the last proof: [0] total, [1] NTT, [2] MSM, [3] quotient,
 * [4] other (CUDA events on the prover's stream); [5] (n >= 6) wall time of the coset NTTs that ran on the prover's second,
 * low-priority stream concurrently with the commitment MSMs (single GPU; not part of [1]). */
int zp_prover_last_timing(zp_prover* p, double* out_ms, int n);

/* ---- ark-serialize 0.3 wire format of `Proof<Fr, KZG10<Bls12_381>>` (proof.rs:37-121) --------------------------
 * 17 compressed G1 commitments (48 B each) || aw_opening (48 B + Option::None byte) || saw_opening (48 B + 1 B) ||
 * wire(4) / permutation(4) / lookup(8) evaluations (32 B LE canonical each) || custom evaluations as
 * Vec<(String, Fr)>: u64 count, then (u64 label length, label bytes, 32 B) x 10.  1930 bytes in total.
 * Host-only helpers (what util.rs:225-291 + CanonicalSerialize do on the Rust side). */
#define ZP_PROOF_SERIALIZED_BYTES 1930
int zp_proof_serialize(const ProofC* proof, uint8_t* out, size_t capacity, size_t* written);
/* Inverse of zp_proof_serialize: decompresses the points (y recovered from x, sign bit) — no subgroup check. */
int zp_proof_deserialize(const uint8_t* bytes, size_t len, ProofC* out);

/* ---- verifier (host side, real pairings) ------------------------------------------------------------------------
 * `Proof::verify` of the reference (proof.rs:123-443; the two `KZG10::check` pairing equations at :414-441;
 * `Verifier::verify`, verifier.rs:106-125) and a batch form that checks any number of proofs of the same circuit with ONE
 * two-pairing product (random linear combination of all opening equations).
 * commitments23: the 19 prover-key polynomial commitments in ProverKeyC order followed by the 4 table commitments (what
 * zp_prover_verifier_key returns; the reference's VerifierKey holds the same points), 12 u64 each.
 * beta_h: [tau] H of the SRS in G2, 24 u64 = x.c0 || x.c1 || y.c0 || y.c1 (Montgomery Fq), ark-poly-commit `vk.beta_h`.
 * Public inputs: positions and Montgomery values (zero values are dropped like `PublicInputs` does, pi.rs:55-62). */
typedef struct zp_verifier zp_verifier;
const char* zp_verifier_last_error(void);
zp_verifier* zp_verifier_create(uint64_t n, const uint64_t* commitments23, const uint64_t* beta_h);
void zp_verifier_destroy(zp_verifier* v);
int zp_verifier_set_label(zp_verifier* v, const char* label);
/* *accepted = 1 iff both opening checks hold; *detail bit 0 = aggregate opening at z, bit 1 = shifted opening at z*omega */
int zp_proof_verify(zp_verifier* v, const ProofC* proof, const uint64_t* pi_pos, const uint64_t* pi_vals, size_t n_pi, int* accepted,
                    int* detail);
/* count proofs, one public input each (pi_pos[i], pi_vals + 4 i) */
int zp_proof_verify_batch(zp_verifier* v, const ProofC* proofs, size_t count, const uint64_t* pi_pos, const uint64_t* pi_vals,
                          int* accepted);
/* scalar * H (test SRS with a known trapdoor: beta_h = tau * H); scalar = Montgomery Fr */
int zp_g2_mul_generator(const uint64_t* scalar, uint64_t* out24);
/* prod_i e(P_i, Q_i): g1_points 12 u64 each, g2_points 24 u64 each; out72 = the GT element (Fq12, Montgomery limbs in
 * tower order c0.c0.c0 .. c1.c2.c1), may be NULL */
int zp_pairing_product(const uint64_t* g1_points, const uint64_t* g2_points, size_t count, uint64_t* out72, int* is_one);

/* ---- operator entry points for the sweeps (function.cuh:45-113 equivalents) ------------------ */
/* kind: 0 NTT, 1 iNTT, 2 coset-NTT (g = 7), 3 coset-iNTT; natural order in/out; host buffers. */
int zp_ntt_host(zp_prover* p, int kind, int log_n, const uint64_t* in, uint64_t* out);
/* Four-step NTT sharded over `world` (1, 2, 4, 8) ranks, one process per GPU: `local_in` / `local_out` are this rank's
 * contiguous blocks [rank*N/world, (rank+1)*N/world) of the natural-order input / output (host memory), N = 2^log_n;
 * `alltoall(user, send_dev, recv_dev, bytes_per_peer)` exchanges DEVICE memory (chunk p of send -> rank p). */
typedef int (*zp_dev_alltoall_fn)(void* user, const void* send_dev, void* recv_dev, size_t bytes_per_peer);
int zp_ntt_sharded_host(zp_prover* p, int kind, int log_n, int rank, int world, const uint64_t* local_in, uint64_t* local_out,
                        zp_dev_alltoall_fn alltoall, void* user);
/* device-resident timing variant: block in slot_in, result in slot_out, two more slots as scratch; *ms = average */
int zp_bench_ntt_sharded(zp_prover* p, int kind, int log_n, int rank, int world, int slot_in, int slot_out, int slot_tmp_a,
                         int slot_tmp_b, int iters, zp_dev_alltoall_fn alltoall, void* user, double* ms);
/* MSM over the first n points of the resident SRS with n host scalars (Montgomery Fr). out: affine. */
int zp_msm_host(zp_prover* p, const uint64_t* scalars, size_t n, uint64_t* out_affine);
/* nbatch (<= 8) MSMs over the same first n SRS points in ONE pipeline: scalars = nbatch * n Montgomery Fr (member-major),
 * out = nbatch affine points.  This is the path gen_proof uses for independent commitments (the four wire
 * polynomials, t_1..t_8, the two opening witnesses; reference: one multi_scalar_mult call each, gen_proof.cuh). */
int zp_msm_batch_host(zp_prover* p, const uint64_t* scalars, int nbatch, size_t n, uint64_t* out_affine);
/* MSM with caller-supplied points (n * 12 u64, host). window_bits = 0 picks the default. */
int zp_msm_points_host(zp_prover* p, const uint64_t* points, const uint64_t* scalars, size_t n, int window_bits,
                       uint64_t* out_affine);
/* p(z) for a host coefficient array */
int zp_poly_eval_host(zp_prover* p, const uint64_t* coeffs, size_t n, const uint64_t* point, uint64_t* out);
/* floor(p / (X - z)) for a host coefficient array (n - 1 coefficients out) */
int zp_poly_divide_host(zp_prover* p, const uint64_t* coeffs, size_t n, const uint64_t* point, uint64_t* out);
/* plookup MultiSet::combine_split(t, f) (lookup/multiset.rs:131-176) on the device: n-element host arrays t, f in,
 * h1, h2 out.  Returns 0, or -1 with "ElementNotIndexed" when an element of f is missing from t. */
int zp_combine_split_host(zp_prover* p, const uint64_t* t, const uint64_t* f, size_t n, uint64_t* h1, uint64_t* h2);
/* The same for multisets of different cardinality (the shape of the reference's own known-answer test,
 * lookup/multiset.rs:335-393): h1 receives ceil((nt + nf) / 2) elements, h2 floor((nt + nf) / 2). */
int zp_multiset_combine_split_host(zp_prover* p, const uint64_t* t, size_t nt, const uint64_t* f, size_t nf, uint64_t* h1,
                                   uint64_t* h2);
/* MultiSet::compress of four columns (lookup/multiset.rs:207-213; `lc`, util.rs:154-176):
 * out[i] = c0[i] + ch c1[i] + ch^2 c2[i] + ch^3 c3[i]; columns = 4 host arrays of n Fr. */
int zp_multiset_compress_host(zp_prover* p, const uint64_t* const* columns, size_t n, const uint64_t* challenge, uint64_t* out);
/* exclusive prefix product (the z(X) scan primitive) */
int zp_prefix_product_host(zp_prover* p, const uint64_t* in, size_t n, uint64_t* out);

/* device-resident variants used by bench.py for kernel-only timing: buffers owned by the context */
int zp_bench_alloc(zp_prover* p, int slot, size_t n_fr);               /* slot 0..7 */
int zp_bench_upload(zp_prover* p, int slot, const uint64_t* host, size_t n_fr);
int zp_bench_download(zp_prover* p, int slot, uint64_t* host, size_t n_fr);
/* runs `iters` transforms slot_in -> slot_out on the stream; returns average device ms via *ms */
int zp_bench_ntt(zp_prover* p, int kind, int log_n, int slot_in, int slot_out, int iters, double* ms);
/* the same for n_in coefficients implicitly zero-padded to 2^log_n (kind 2 with log_n = logN + 3, n_in = N is the
 * extended-coset transform of the quotient round; reference: Ntt_coset::forward, function.cu:261-268) */
int zp_bench_ntt_padded(zp_prover* p, int kind, int log_n, size_t n_in, int slot_in, int slot_out, int iters, double* ms);
/* runs `iters` MSMs of n points with scalars in `slot`; *ms = average device ms (host tail included) */
int zp_bench_msm(zp_prover* p, int slot, size_t n, int iters, double* ms, uint64_t* out_affine);
/* the same for `nbatch` (<= 8) scalar vectors over the same points in ONE pipeline (the batch the prover uses for
 * independent commitments); *ms = average device ms per batch */
int zp_bench_msm_batch(zp_prover* p, int slot, size_t n, int nbatch, int iters, double* ms, uint64_t* out_affine);
/* the same through the prover's commitment path, which honours zp_prover_set_shard: every rank calls it with the same scalars
 * (one process per GPU), the MSM is split across the ranks (bucket shares with precomputed tables, point ranges below
 * 2^16) and the partial sums are all-gathered; *ms = average device ms per batch on this rank, out_affine = the full sum */
int zp_bench_commit_sharded(zp_prover* p, int slot, size_t n, int nbatch, int iters, double* ms, uint64_t* out_affine);
/* per-stage device time of the last zp_bench_msm iteration: digits, scan, scatter, batch-affine rounds,
 * accumulate (+ folds), reduce */
int zp_bench_msm_breakdown(zp_prover* p, double* ms6);
/* pipe microbenchmarks: mode 0 = IMAD (mad.lo), 1 = IMAD.WIDE (mad.wide), 2 = Fq Montgomery products,
 * 3 = Fq Montgomery squarings, 4 = FP64 FMA, 5 = FP64 FMA interleaved 1:1 with IMAD (FP64 operations counted),
 * 6 = 3-input integer add (ALU pipe); returns giga-operations per second (instructions, or field operations for 2/3) */
int zp_bench_int_pipe(zp_prover* p, int mode, double* gops);

#ifdef __cplusplus
}
#endif
#endif /* ZPRIZE_B200_H */
