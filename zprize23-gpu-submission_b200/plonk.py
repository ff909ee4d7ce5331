"""Host-side mirror of the reference's FFI surface for the PLONK prover hot path.

The reference's host side is Rust (`Prover::prove_pnp`, "Prize 1B/plonk-core/src/proof_system/prover.rs":693-907,
binding `extern "C" gen_proof` declared at "Prize 1B/plonk-core/src/lib.rs":237-239).  No Rust toolchain exists in
this image, so the marshalling layer above the C-ABI is mirrored here with ctypes: the same `#[repr(C)]` structs
(lib.rs:53-235), the same `gen_proof(CircuitC, ProverKeyC, CommitKeyC) -> ProofC` call, plus the resident-context
extension of include/zprize_b200.h.  All numerical work happens in libzprize_b200.so (hand-written sm_100a CUDA);
there is no CPU path: loading fails loudly if the library is missing, and context creation fails without a GPU.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzprize_b200.so")

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)


class CommitmentC(ctypes.Structure):  # lib.rs:231-235
    _fields_ = [("x", ctypes.c_uint64 * 6), ("y", ctypes.c_uint64 * 6)]


def _fr_fields(names):
    return [(n, ctypes.c_uint64 * 4) for n in names]


class WireEvaluationsC(ctypes.Structure):  # lib.rs:53-59
    _fields_ = _fr_fields(["a_eval", "b_eval", "c_eval", "d_eval"])


class PermutationEvaluationsC(ctypes.Structure):  # lib.rs:61-67
    _fields_ = _fr_fields(["left_sigma_eval", "right_sigma_eval", "out_sigma_eval", "permutation_eval"])


class LookupEvaluationsC(ctypes.Structure):  # lib.rs:100-110
    _fields_ = _fr_fields(["q_lookup_eval", "z2_next_eval", "h1_eval", "h1_next_eval", "h2_eval", "f_eval", "table_eval",
                           "table_next_eval"])


class CustomEvaluationsC(ctypes.Structure):  # lib.rs:69-81
    _fields_ = _fr_fields(["q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval", "q_hr_eval", "q_h4_eval",
                           "a_next_eval", "b_next_eval", "d_next_eval"])


class ProofEvaluationsC(ctypes.Structure):  # lib.rs:111-118
    _fields_ = [("wire_evals", WireEvaluationsC), ("perm_evals", PermutationEvaluationsC),
                ("lookup_evals", LookupEvaluationsC), ("custom_evals", CustomEvaluationsC)]


COMMITMENT_NAMES = ["a_comm", "b_comm", "c_comm", "d_comm", "z_comm", "f_comm", "h_1_comm", "h_2_comm", "z_2_comm",
                    "t_1_comm", "t_2_comm", "t_3_comm", "t_4_comm", "t_5_comm", "t_6_comm", "t_7_comm", "t_8_comm",
                    "aw_opening", "saw_opening"]
EVALUATION_NAMES = (["a_eval", "b_eval", "c_eval", "d_eval", "left_sigma_eval", "right_sigma_eval", "out_sigma_eval",
                     "permutation_eval", "q_lookup_eval", "z2_next_eval", "h1_eval", "h1_next_eval", "h2_eval", "f_eval",
                     "table_eval", "table_next_eval", "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "q_hl_eval",
                     "q_hr_eval", "q_h4_eval", "a_next_eval", "b_next_eval", "d_next_eval"])


class ProofC(ctypes.Structure):  # lib.rs:120-142
    _fields_ = [(n, CommitmentC) for n in COMMITMENT_NAMES] + [("evaluations", ProofEvaluationsC)]

    def to_words(self):
        """The 2656-byte image as 332 little-endian u64 words."""
        return np.frombuffer(bytes(self), dtype=np.uint64).copy()


class CircuitC(ctypes.Structure):  # lib.rs:144-155
    _fields_ = [("n", ctypes.c_uint64), ("lookup_len", ctypes.c_uint64), ("intended_pi_pos", ctypes.c_uint64),
                ("q_lookup", u64p), ("pi", u64p), ("w_l", u64p), ("w_r", u64p), ("w_o", u64p), ("w_4", u64p)]


PK_POLY_NAMES = ["q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_hl", "q_hr", "q_h4", "q_arith", "range_selector",
                 "logic_selector", "fixed_group_add_selector", "variable_group_add_selector", "q_lookup"]
PK_SIGMA_NAMES = ["left_sigma", "right_sigma", "out_sigma", "fourth_sigma"]


def _pk_fields():
    f = []
    for n in PK_POLY_NAMES:
        f += [(n + "_coeffs", u64p), (n + "_evals", u64p)]
    f += [("table1", u64p), ("table2", u64p), ("table3", u64p), ("table4", u64p)]
    for n in PK_SIGMA_NAMES:
        f += [(n + "_coeffs", u64p), (n + "_evals", u64p)]
    f += [("linear_evaluations", u64p), ("v_h_coset_8n", u64p)]
    return f


class ProverKeyC(ctypes.Structure):  # lib.rs:157-223
    _fields_ = _pk_fields()


class CommitKeyC(ctypes.Structure):  # lib.rs:225-229
    _fields_ = [("powers_of_g", u64p), ("powers_of_gamma_g", u64p)]


assert ctypes.sizeof(ProofC) == 2656 and ctypes.sizeof(ProverKeyC) == 44 * 8 and ctypes.sizeof(CircuitC) == 72


# int (*zp_allgather_fn)(void* user, const void* send, void* recv, size_t bytes_per_rank)
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


# int (*zp_dev_broadcast_fn)(void* user, void* dev_ptr, size_t bytes, int root)
DEV_BCAST_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int)


# int (*zp_dev_allgather_fn)(void* user, void* dev_base, size_t bytes_per_rank)
DEV_ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


# int (*zp_dev_alltoall_fn)(void* user, const void* send_dev, void* recv_dev, size_t bytes_per_peer)
DEV_A2A_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


class ZprizeError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """Loads libzprize_b200.so (built in-tree by build.py).  No fallback: a missing library is an error."""
    global _lib
    path = path or LIB_PATH
    if _lib is not None and getattr(_lib, "_zp_path", None) == path:
        return _lib
    if not os.path.exists(path):
        raise ZprizeError("%s not found: build it with `python %s` (nvcc, sm_100a); there is no CPU fallback"
                          % (path, os.path.join(_HERE, "build.py")))
    lib = ctypes.CDLL(path)
    lib._zp_path = path
    vp, ci, cs, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_double
    dp = ctypes.POINTER(ctypes.c_double)
    sig = {
        "zp_last_error": (ctypes.c_char_p, []),
        "zp_launch_count": (ctypes.c_uint64, []),
        "zp_device_available": (ci, []),
        "zp_prover_create": (vp, [ci]),
        "zp_prover_destroy": (None, [vp]),
        "zp_prover_set_label": (ci, [vp, ctypes.c_char_p]),
        "zp_profiler_range": (ci, [ci]),
        "zp_prover_set_stream": (ci, [vp, vp]),
        "zp_prover_load_srs": (ci, [vp, u64p, cs]),
        "zp_prover_generate_srs": (ci, [vp, u64p, cs]),
        "zp_prover_read_srs": (ci, [vp, u64p, cs]),
        "zp_prover_load_pk": (ci, [vp, ctypes.POINTER(ProverKeyC), u64p]),
        "zp_prover_preprocess": (ci, [vp, ctypes.POINTER(u64p), ctypes.POINTER(u64p)]),
        "zp_prover_read_pk": (ci, [vp, ci, u64p, u64p]),
        "zp_prover_preprocess_wiring": (ci, [vp, ctypes.POINTER(u64p), u32p, u32p, cs, ctypes.c_uint32, ctypes.POINTER(u64p)]),
        "zp_sigma_from_wiring_host": (ci, [vp, u32p, u32p, cs, ctypes.c_uint32, ctypes.POINTER(u64p)]),
        "zp_prover_verifier_key": (ci, [vp, u64p]),
        "zp_prover_prove": (ci, [vp, ctypes.POINTER(CircuitC), ctypes.POINTER(ProofC)]),
        "zp_prover_last_timing": (ci, [vp, dp, ci]),
        "zp_prover_upload_witness": (ci, [vp, ctypes.POINTER(CircuitC)]),
        "zp_prover_prove_resident": (ci, [vp, ctypes.POINTER(ProofC)]),
        "zp_prover_synthesize_merkle_witness": (ci, [vp, ci, u64p, u64p, u64p, u64p]),
        "zp_prover_read_witness": (ci, [vp, ci, u64p]),
        "zp_prover_witness_rows": (ctypes.c_uint64, [vp]),
        "zp_prover_collect_msm_stats": (ci, [vp, ci]),
        "zp_prover_msm_stats": (ci, [vp, dp]),
        "zp_prover_set_shard": (ci, [vp, ci, ci, ALLGATHER_FN, vp]),
        "zp_prover_set_device_broadcast": (ci, [vp, DEV_BCAST_FN, vp]),
        "zp_prover_set_device_allgather": (ci, [vp, DEV_ALLGATHER_FN, vp]),
        "zp_ntt_host": (ci, [vp, ci, ci, u64p, u64p]),
        "zp_msm_host": (ci, [vp, u64p, cs, u64p]),
        "zp_msm_batch_host": (ci, [vp, u64p, ci, cs, u64p]),
        "zp_ntt_sharded_host": (ci, [vp, ci, ci, ci, ci, u64p, u64p, DEV_A2A_FN, vp]),
        "zp_bench_ntt_sharded": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, DEV_A2A_FN, vp, dp]),
        "zp_msm_points_host": (ci, [vp, u64p, u64p, cs, ci, u64p]),
        "zp_poly_eval_host": (ci, [vp, u64p, cs, u64p, u64p]),
        "zp_poly_divide_host": (ci, [vp, u64p, cs, u64p, u64p]),
        "zp_prefix_product_host": (ci, [vp, u64p, cs, u64p]),
        "zp_combine_split_host": (ci, [vp, u64p, u64p, cs, u64p, u64p]),
        "zp_multiset_combine_split_host": (ci, [vp, u64p, cs, u64p, cs, u64p, u64p]),
        "zp_multiset_compress_host": (ci, [vp, ctypes.POINTER(u64p), cs, u64p, u64p]),
        "zp_bench_alloc": (ci, [vp, ci, cs]),
        "zp_bench_upload": (ci, [vp, ci, u64p, cs]),
        "zp_bench_download": (ci, [vp, ci, u64p, cs]),
        "zp_bench_ntt": (ci, [vp, ci, ci, ci, ci, ci, dp]),
        "zp_bench_ntt_padded": (ci, [vp, ci, ci, cs, ci, ci, ci, dp]),
        "zp_bench_msm": (ci, [vp, ci, cs, ci, dp, u64p]),
        "zp_bench_msm_batch": (ci, [vp, ci, cs, ci, ci, dp, u64p]),
        "zp_bench_commit_sharded": (ci, [vp, ci, cs, ci, ci, dp, u64p]),
        "zp_bench_msm_breakdown": (ci, [vp, dp]),
        "zp_bench_int_pipe": (ci, [vp, ci, dp]),
        "zp_proof_serialize": (ci, [ctypes.POINTER(ProofC), ctypes.c_char_p, cs, ctypes.POINTER(cs)]),
        "zp_proof_deserialize": (ci, [ctypes.c_char_p, cs, ctypes.POINTER(ProofC)]),
        "zp_verifier_last_error": (ctypes.c_char_p, []),
        "zp_verifier_create": (vp, [ctypes.c_uint64, u64p, u64p]),
        "zp_verifier_destroy": (None, [vp]),
        "zp_verifier_set_label": (ci, [vp, ctypes.c_char_p]),
        "zp_proof_verify": (ci, [vp, ctypes.POINTER(ProofC), u64p, u64p, cs, ctypes.POINTER(ci), ctypes.POINTER(ci)]),
        "zp_proof_verify_batch": (ci, [vp, ctypes.POINTER(ProofC), cs, u64p, u64p, ctypes.POINTER(ci)]),
        "zp_g2_mul_generator": (ci, [u64p, u64p]),
        "zp_pairing_product": (ci, [u64p, u64p, cs, u64p, ctypes.POINTER(ci)]),
        "gen_proof": (ProofC, [CircuitC, ProverKeyC, CommitKeyC]),
        "zp_gen_proof_invalidate": (None, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ["gen_proof", "zp_verifier_last_error", "zp_verifier_create", "zp_verifier_destroy", "zp_verifier_set_label", "zp_proof_verify", "zp_proof_verify_batch", "zp_g2_mul_generator", "zp_pairing_product", "zp_gen_proof_invalidate", "zp_proof_serialize", "zp_proof_deserialize", "zp_last_error", "zp_launch_count", "zp_device_available", "zp_prover_create",
                    "zp_prover_destroy", "zp_prover_set_label", "zp_profiler_range", "zp_prover_set_stream", "zp_prover_load_srs", "zp_prover_generate_srs",
                    "zp_prover_read_srs", "zp_prover_load_pk", "zp_prover_preprocess", "zp_prover_read_pk", "zp_prover_preprocess_wiring", "zp_sigma_from_wiring_host", "zp_prover_verifier_key",
                    "zp_prover_prove", "zp_prover_last_timing", "zp_prover_upload_witness", "zp_prover_prove_resident", "zp_prover_synthesize_merkle_witness", "zp_prover_read_witness", "zp_prover_witness_rows",
                    "zp_prover_collect_msm_stats", "zp_prover_msm_stats", "zp_prover_set_shard", "zp_prover_set_device_broadcast", "zp_prover_set_device_allgather", "zp_ntt_host", "zp_ntt_sharded_host", "zp_bench_ntt_sharded", "zp_msm_host", "zp_msm_batch_host", "zp_msm_points_host",
                    "zp_poly_eval_host", "zp_poly_divide_host", "zp_prefix_product_host", "zp_combine_split_host", "zp_multiset_combine_split_host", "zp_multiset_compress_host", "zp_bench_alloc",
                    "zp_bench_upload", "zp_bench_download", "zp_bench_ntt", "zp_bench_ntt_padded", "zp_bench_msm", "zp_bench_msm_batch", "zp_bench_commit_sharded", "zp_bench_msm_breakdown",
                    "zp_bench_int_pipe"]


def as_u64p(a):
    """Pointer to a C-contiguous uint64 numpy array (kept alive by the caller)."""
    if a is None:
        return ctypes.cast(None, u64p)
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def make_circuit(n, lookup_len, pi_pos, q_lookup, pi_canonical, w_l, w_r, w_o, w_4):
    """CircuitC over numpy uint64 arrays ([n,4] Fr words each; pi: 4 words, canonical form)."""
    c = CircuitC()
    c.n, c.lookup_len, c.intended_pi_pos = n, lookup_len, pi_pos
    c.q_lookup, c.pi = as_u64p(q_lookup), as_u64p(pi_canonical)
    c.w_l, c.w_r, c.w_o, c.w_4 = as_u64p(w_l), as_u64p(w_r), as_u64p(w_o), as_u64p(w_4)
    c._keep = (q_lookup, pi_canonical, w_l, w_r, w_o, w_4)
    return c


def make_prover_key(coeffs, evals, tables, linear_evaluations=None, v_h_coset_8n=None):
    """ProverKeyC from dicts name -> array (names of PK_POLY_NAMES + PK_SIGMA_NAMES) and 4 table columns."""
    pk = ProverKeyC()
    keep = []
    for nme in PK_POLY_NAMES + PK_SIGMA_NAMES:
        setattr(pk, nme + "_coeffs", as_u64p(coeffs.get(nme)))
        setattr(pk, nme + "_evals", as_u64p(evals.get(nme)))
        keep += [coeffs.get(nme), evals.get(nme)]
    for i in range(4):
        setattr(pk, "table%d" % (i + 1), as_u64p(tables[i]))
    pk.linear_evaluations = as_u64p(linear_evaluations)
    pk.v_h_coset_8n = as_u64p(v_h_coset_8n)
    pk._keep = (keep, tables, linear_evaluations, v_h_coset_8n)
    return pk


PROOF_SERIALIZED_BYTES = 1930


def proof_serialize(proof, lib=None):
    """ark-serialize 0.3 bytes of `Proof<Fr, KZG10<Bls12_381>>` (what `CanonicalSerialize` gives the Rust side)."""
    lib = lib or load_library()
    buf = ctypes.create_string_buffer(PROOF_SERIALIZED_BYTES)
    n = ctypes.c_size_t()
    if lib.zp_proof_serialize(ctypes.byref(proof), buf, PROOF_SERIALIZED_BYTES, ctypes.byref(n)) != 0:
        raise ZprizeError(lib.zp_last_error().decode())
    return buf.raw[:n.value]


def proof_deserialize(data, lib=None):
    lib = lib or load_library()
    proof = ProofC()
    if lib.zp_proof_deserialize(data, len(data), ctypes.byref(proof)) != 0:
        raise ZprizeError(lib.zp_last_error().decode())
    return proof


def gen_proof(circuit, pk, ck, lib=None):
    """The reference's FFI call itself: `gen_proof(CircuitC, ProverKeyC, CommitKeyC) -> ProofC` (by value)."""
    lib = lib or load_library()
    return lib.gen_proof(circuit, pk, ck)


def proof_from_words(words):
    """ProofC from its 2656-byte image (332 u64 words)."""
    return ProofC.from_buffer_copy(np.ascontiguousarray(words, dtype=np.uint64).tobytes())


def g2_mul_generator(scalar_words, lib=None):
    lib = lib or load_library()
    out = np.zeros(24, dtype=np.uint64)
    if lib.zp_g2_mul_generator(as_u64p(np.ascontiguousarray(scalar_words, dtype=np.uint64)), as_u64p(out)) != 0:
        raise ZprizeError(lib.zp_verifier_last_error().decode())
    return out


def pairing_product(g1_points, g2_points, lib=None):
    """(GT element as 72 u64 words, is_one) of prod e(P_i, Q_i)."""
    lib = lib or load_library()
    g1 = np.ascontiguousarray(g1_points, dtype=np.uint64).reshape(-1, 12)
    g2 = np.ascontiguousarray(g2_points, dtype=np.uint64).reshape(-1, 24)
    out = np.zeros(72, dtype=np.uint64)
    one = ctypes.c_int()
    if lib.zp_pairing_product(as_u64p(g1), as_u64p(g2), g1.shape[0], as_u64p(out), ctypes.byref(one)) != 0:
        raise ZprizeError(lib.zp_verifier_last_error().decode())
    return out, one.value == 1


class Verifier:
    """`Proof::verify` with real pairings (host side of the library; needs no GPU)."""

    def __init__(self, n, commitments23, beta_h, lib=None, label=None):
        self.lib = lib or load_library()
        c = np.ascontiguousarray(commitments23, dtype=np.uint64)
        b = np.ascontiguousarray(beta_h, dtype=np.uint64)
        self.h = self.lib.zp_verifier_create(n, as_u64p(c), as_u64p(b))
        if not self.h:
            raise ZprizeError(self.lib.zp_verifier_last_error().decode())
        if label is not None and self.lib.zp_verifier_set_label(self.h, label) != 0:
            raise ZprizeError(self.lib.zp_verifier_last_error().decode())

    def close(self):
        if self.h:
            self.lib.zp_verifier_destroy(self.h)
            self.h = None

    def verify(self, proof, pi_pos, pi_val_mont):
        """(accepted, detail): detail bit 0 = opening at z, bit 1 = opening at z * omega.  pi_val_mont None = no public input."""
        if not isinstance(proof, ProofC):
            proof = proof_from_words(proof)
        acc, det = ctypes.c_int(), ctypes.c_int()
        if pi_val_mont is None:
            pos, val, cnt = np.zeros(1, np.uint64), np.zeros(4, np.uint64), 0
        else:
            pos, val, cnt = np.array([pi_pos], dtype=np.uint64), np.ascontiguousarray(pi_val_mont, dtype=np.uint64), 1
        if self.lib.zp_proof_verify(self.h, ctypes.byref(proof), as_u64p(pos), as_u64p(val), cnt, ctypes.byref(acc),
                                    ctypes.byref(det)) != 0:
            raise ZprizeError(self.lib.zp_verifier_last_error().decode())
        return acc.value == 1, det.value

    def verify_batch(self, proofs, pi_pos, pi_vals_mont):
        arr = (ProofC * len(proofs))(*[p if isinstance(p, ProofC) else proof_from_words(p) for p in proofs])
        pos = np.ascontiguousarray(pi_pos, dtype=np.uint64)
        val = np.ascontiguousarray(pi_vals_mont, dtype=np.uint64)
        acc = ctypes.c_int()
        if self.lib.zp_proof_verify_batch(self.h, arr, len(proofs), as_u64p(pos), as_u64p(val), ctypes.byref(acc)) != 0:
            raise ZprizeError(self.lib.zp_verifier_last_error().decode())
        return acc.value == 1


class ProverContext:
    """Resident prover (extension of the FFI): SRS, prover key, twiddles and work buffers stay in HBM."""

    def __init__(self, log_n, lib=None):
        self.lib = lib or load_library()
        self.log_n = log_n
        self.n = 1 << log_n
        self.h = self.lib.zp_prover_create(log_n)
        if not self.h:
            raise ZprizeError(self.lib.zp_last_error().decode())

    def close(self):
        if self.h:
            self.lib.zp_prover_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise ZprizeError(self.lib.zp_last_error().decode())

    def set_stream(self, cuda_stream):
        self._ck(self.lib.zp_prover_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def set_label(self, label):
        self._ck(self.lib.zp_prover_set_label(self.h, label))

    def load_srs(self, points):
        self._ck(self.lib.zp_prover_load_srs(self.h, as_u64p(points), points.size // 12))

    def generate_srs(self, tau_words, n_points=None):
        self._ck(self.lib.zp_prover_generate_srs(self.h, as_u64p(tau_words), n_points or self.n))

    def read_srs(self, n_points=None):
        n_points = n_points or self.n
        out = np.zeros((n_points, 12), dtype=np.uint64)
        self._ck(self.lib.zp_prover_read_srs(self.h, as_u64p(out), n_points))
        return out

    def load_pk(self, pk, coeff_len=None):
        cl = None if coeff_len is None else np.asarray(coeff_len, dtype=np.uint64)
        self._ck(self.lib.zp_prover_load_pk(self.h, ctypes.byref(pk), as_u64p(cl)))

    def preprocess(self, selector_evals, tables=None):
        """selector_evals: 19 arrays (15 selectors + 4 sigmas) of N Fr on H, or None for all-zero."""
        arr = (u64p * 19)(*[as_u64p(a) for a in selector_evals])
        tb = None
        if tables is not None:
            tb = (u64p * 4)(*[as_u64p(a) for a in tables])
        self._ck(self.lib.zp_prover_preprocess(self.h, arr, tb))

    def preprocess_wiring(self, selector_evals15, wire_vars, wire_cells, n_vars, tables=None):
        """Preprocessing with the sigma polynomials built on the device from the wire map (uint32 arrays of equal length:
        variable id and (gate << 2) | wire per entry, in the composer's insertion order)."""
        arr = (u64p * 15)(*[as_u64p(a) for a in selector_evals15])
        tb = None
        if tables is not None:
            tb = (u64p * 4)(*[as_u64p(a) for a in tables])
        v = np.ascontiguousarray(wire_vars, dtype=np.uint32)
        c = np.ascontiguousarray(wire_cells, dtype=np.uint32)
        self._ck(self.lib.zp_prover_preprocess_wiring(self.h, arr, v.ctypes.data_as(u32p), c.ctypes.data_as(u32p), v.shape[0],
                                                      n_vars, tb))

    def sigma_from_wiring(self, wire_vars, wire_cells, n_vars):
        v = np.ascontiguousarray(wire_vars, dtype=np.uint32)
        c = np.ascontiguousarray(wire_cells, dtype=np.uint32)
        out = [np.zeros((self.n, 4), dtype=np.uint64) for _ in range(4)]
        arr = (u64p * 4)(*[as_u64p(a) for a in out])
        self._ck(self.lib.zp_sigma_from_wiring_host(self.h, v.ctypes.data_as(u32p), c.ctypes.data_as(u32p), v.shape[0], n_vars, arr))
        return out

    def read_pk(self, index, want_coeffs=True, want_evals=True):
        """(coeffs [N,4] or None, evals [8N,4] or None) of prover-key polynomial `index` (ProverKeyC order; 19..22 = tables)."""
        co = np.zeros((self.n, 4), dtype=np.uint64) if want_coeffs else None
        ev = np.zeros((8 * self.n, 4), dtype=np.uint64) if (want_evals and index < 19) else None
        self._ck(self.lib.zp_prover_read_pk(self.h, index, as_u64p(co), as_u64p(ev)))
        return co, ev

    def verifier_key(self):
        out = np.zeros((23, 12), dtype=np.uint64)
        self._ck(self.lib.zp_prover_verifier_key(self.h, as_u64p(out)))
        return out

    def prove(self, circuit):
        proof = ProofC()
        self._ck(self.lib.zp_prover_prove(self.h, ctypes.byref(circuit), ctypes.byref(proof)))
        return proof

    def upload_witness(self, circuit):
        self._ck(self.lib.zp_prover_upload_witness(self.h, ctypes.byref(circuit)))

    def synthesize_merkle_witness(self, height, leaves, hash_params, blinding):
        """Builds the Poseidon-Merkle witness on the device (resident for prove_resident); returns the root (Montgomery)."""
        root = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.zp_prover_synthesize_merkle_witness(self.h, height, as_u64p(leaves), as_u64p(hash_params),
                                                              as_u64p(blinding), as_u64p(root)))
        return root

    def read_witness(self):
        rows = self.lib.zp_prover_witness_rows(self.h)
        out = [np.zeros((rows, 4), dtype=np.uint64) for _ in range(4)]
        for k in range(4):
            self._ck(self.lib.zp_prover_read_witness(self.h, k, as_u64p(out[k])))
        return out

    def prove_resident(self):
        proof = ProofC()
        self._ck(self.lib.zp_prover_prove_resident(self.h, ctypes.byref(proof)))
        return proof

    def collect_msm_stats(self, enable=True):
        self._ck(self.lib.zp_prover_collect_msm_stats(self.h, 1 if enable else 0))

    def msm_stats(self):
        out = (ctypes.c_double * 9)()
        self._ck(self.lib.zp_prover_msm_stats(self.h, out))
        return {"accumulate_ms": out[0], "launches": int(out[1]), "algorithmic_mads": out[2], "all_kernels_ms": out[3],
                "executed_mads": out[4], "commitments": int(out[5]), "down0_ms": out[6], "down0_pairs": out[7],
                "down0_launches": int(out[8])}

    def set_shard(self, rank, world, allgather):
        """allgather(send_bytes: bytes) -> bytes of world * len(send_bytes) in rank order (e.g. torch.distributed)."""
        def _cb(user, send, recv, nbytes):
            try:
                data = ctypes.string_at(send, nbytes)
                out = allgather(data)
                assert len(out) == nbytes * world
                ctypes.memmove(recv, out, len(out))
                return 0
            except Exception as e:  # noqa: BLE001 - reported through the C error channel
                self._cb_error = e
                return 1
        self._cb = ALLGATHER_FN(_cb) if world > 1 else ctypes.cast(None, ALLGATHER_FN)
        self._ck(self.lib.zp_prover_set_shard(self.h, rank, world, self._cb, None))

    def set_device_broadcast(self, bcast):
        """bcast(dev_ptr: int, nbytes: int, root: int) broadcasts device memory in place (None disables)."""
        if bcast is None:
            self._bc = ctypes.cast(None, DEV_BCAST_FN)
        else:
            def _cb(user, ptr, nbytes, root):
                try:
                    bcast(ptr, nbytes, root)
                    return 0
                except Exception as e:  # noqa: BLE001
                    self._cb_error = e
                    return 1
            self._bc = DEV_BCAST_FN(_cb)
        self._ck(self.lib.zp_prover_set_device_broadcast(self.h, self._bc, None))

    def set_device_allgather(self, allgather):
        """allgather(dev_base: int, bytes_per_rank: int): in-place all-gather of device memory — rank r's block
        [r * bytes_per_rank, (r + 1) * bytes_per_rank) reaches every rank (None disables)."""
        if allgather is None:
            self._ag = ctypes.cast(None, DEV_ALLGATHER_FN)
        else:
            def _cb(user, ptr, nbytes):
                try:
                    allgather(ptr, nbytes)
                    return 0
                except Exception as e:  # noqa: BLE001
                    self._cb_error = e
                    return 1
            self._ag = DEV_ALLGATHER_FN(_cb)
        self._ck(self.lib.zp_prover_set_device_allgather(self.h, self._ag, None))

    def last_timing(self):
        out = (ctypes.c_double * 6)()
        self._ck(self.lib.zp_prover_last_timing(self.h, out, 6))
        return dict(zip(["total_ms", "ntt_ms", "msm_ms", "quotient_ms", "other_ms", "ntt_overlapped_ms"], list(out)))

    # ---- operator entry points (function.cuh:45-113 equivalents) ----
    def ntt(self, kind, data):
        log_n = int(np.log2(data.shape[0]))
        out = np.zeros_like(data)
        self._ck(self.lib.zp_ntt_host(self.h, kind, log_n, as_u64p(data), as_u64p(out)))
        return out

    def _a2a_cb(self, alltoall):
        def _cb(user, send, recv, nbytes):
            try:
                alltoall(send, recv, nbytes)
                return 0
            except Exception as e:  # noqa: BLE001
                self._cb_error = e
                return 1
        return DEV_A2A_FN(_cb)

    def ntt_sharded(self, kind, log_n, rank, world, local_in, alltoall):
        """alltoall(send_ptr: int, recv_ptr: int, bytes_per_peer: int) exchanges device memory between the ranks."""
        out = np.zeros_like(local_in)
        cb = self._a2a_cb(alltoall)
        self._ck(self.lib.zp_ntt_sharded_host(self.h, kind, log_n, rank, world, as_u64p(local_in), as_u64p(out), cb, None))
        return out

    def bench_ntt_sharded(self, kind, log_n, rank, world, slots, iters, alltoall):
        ms = ctypes.c_double()
        cb = self._a2a_cb(alltoall)
        self._ck(self.lib.zp_bench_ntt_sharded(self.h, kind, log_n, rank, world, slots[0], slots[1], slots[2], slots[3], iters, cb,
                                               None, ctypes.byref(ms)))
        return ms.value

    def msm(self, scalars):
        out = np.zeros(12, dtype=np.uint64)
        self._ck(self.lib.zp_msm_host(self.h, as_u64p(scalars), scalars.shape[0], as_u64p(out)))
        return out

    def msm_batch(self, scalars):
        """scalars: (k, n, 4) uint64, k <= 8 -> (k, 12) affine points; one MSM pipeline for all k members."""
        scalars = np.ascontiguousarray(scalars)
        k, n = scalars.shape[0], scalars.shape[1]
        out = np.zeros((k, 12), dtype=np.uint64)
        self._ck(self.lib.zp_msm_batch_host(self.h, as_u64p(scalars), k, n, as_u64p(out)))
        return out

    def msm_points(self, points, scalars, window_bits=0):
        out = np.zeros(12, dtype=np.uint64)
        self._ck(self.lib.zp_msm_points_host(self.h, as_u64p(points), as_u64p(scalars), scalars.shape[0], window_bits,
                                             as_u64p(out)))
        return out

    def poly_eval(self, coeffs, point):
        out = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.zp_poly_eval_host(self.h, as_u64p(coeffs), coeffs.shape[0], as_u64p(point), as_u64p(out)))
        return out

    def poly_divide(self, coeffs, point):
        out = np.zeros((coeffs.shape[0] - 1, 4), dtype=np.uint64)
        self._ck(self.lib.zp_poly_divide_host(self.h, as_u64p(coeffs), coeffs.shape[0], as_u64p(point), as_u64p(out)))
        return out

    def combine_split(self, t, f):
        h1, h2 = np.zeros_like(t), np.zeros_like(t)
        self._ck(self.lib.zp_combine_split_host(self.h, as_u64p(t), as_u64p(f), t.shape[0], as_u64p(h1), as_u64p(h2)))
        return h1, h2

    def multiset_combine_split(self, t, f):
        """MultiSet::combine_split for |t| != |f| (lookup/multiset.rs:131-176)."""
        nt, nf = t.shape[0], f.shape[0]
        h1 = np.zeros(((nt + nf + 1) // 2, 4), dtype=np.uint64)
        h2 = np.zeros(((nt + nf) // 2, 4), dtype=np.uint64)
        self._ck(self.lib.zp_multiset_combine_split_host(self.h, as_u64p(t), nt, as_u64p(f), nf, as_u64p(h1), as_u64p(h2)))
        return h1, h2

    def multiset_compress(self, columns, challenge):
        """MultiSet::compress (lookup/multiset.rs:207-213) of four columns with challenge `challenge`."""
        arr = (u64p * 4)(*[as_u64p(c) for c in columns])
        out = np.zeros_like(columns[0])
        self._ck(self.lib.zp_multiset_compress_host(self.h, arr, columns[0].shape[0], as_u64p(challenge), as_u64p(out)))
        return out

    def prefix_product(self, data):
        out = np.zeros_like(data)
        self._ck(self.lib.zp_prefix_product_host(self.h, as_u64p(data), data.shape[0], as_u64p(out)))
        return out

    # ---- device-resident benchmark helpers ----
    def bench_alloc(self, slot, n_fr):
        self._ck(self.lib.zp_bench_alloc(self.h, slot, n_fr))

    def bench_upload(self, slot, data):
        self._ck(self.lib.zp_bench_upload(self.h, slot, as_u64p(data), data.size // 4))

    def bench_download(self, slot, n_fr):
        out = np.zeros((n_fr, 4), dtype=np.uint64)
        self._ck(self.lib.zp_bench_download(self.h, slot, as_u64p(out), n_fr))
        return out

    def bench_ntt(self, kind, log_n, slot_in, slot_out, iters):
        ms = ctypes.c_double()
        self._ck(self.lib.zp_bench_ntt(self.h, kind, log_n, slot_in, slot_out, iters, ctypes.byref(ms)))
        return ms.value

    def bench_msm(self, slot, n, iters, nbatch=1):
        ms = ctypes.c_double()
        out = np.zeros(12, dtype=np.uint64)
        self._ck(self.lib.zp_bench_msm_batch(self.h, slot, n, nbatch, iters, ctypes.byref(ms), as_u64p(out)))
        bd = (ctypes.c_double * 6)()
        self._ck(self.lib.zp_bench_msm_breakdown(self.h, bd))
        return ms.value, out, dict(zip(["digits", "scan", "scatter", "batch_affine", "accumulate", "reduce"], list(bd)))

    def bench_commit_sharded(self, slot, n, iters, nbatch=1):
        ms = ctypes.c_double()
        out = np.zeros(12, dtype=np.uint64)
        self._ck(self.lib.zp_bench_commit_sharded(self.h, slot, n, nbatch, iters, ctypes.byref(ms), as_u64p(out)))
        return ms.value, out

    def bench_int_pipe(self, mode):
        g = ctypes.c_double()
        self._ck(self.lib.zp_bench_int_pipe(self.h, mode, ctypes.byref(g)))
        return g.value
