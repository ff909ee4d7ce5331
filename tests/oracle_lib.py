"""ctypes wrapper around oracle/liboracle.so (CPU restatement — test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
u64p = ctypes.POINTER(ctypes.c_uint64)

PK_NAMES = ["q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_hl", "q_hr", "q_h4", "q_arith", "range_selector",
            "logic_selector", "fixed_group_add_selector", "variable_group_add_selector", "q_lookup", "left_sigma",
            "right_sigma", "out_sigma", "fourth_sigma"]


def _p(a):
    if a is None:
        return ctypes.cast(None, u64p)
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        vp, ci, cs, c64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64
        L.zpo_ctx_new.restype = vp
        L.zpo_ctx_new.argtypes = [ci, c64, c64, ci, ci, ci]
        L.zpo_ctx_new_kind.restype = vp
        L.zpo_ctx_new_kind.argtypes = [ci, ci, c64, c64, ci, ci, ci]
        L.zpo_ctx_free.argtypes = [vp]
        L.zpo_ctx_wiring.argtypes = [vp, ctypes.c_void_p, ctypes.c_void_p]
        for name in ["zpo_ctx_n", "zpo_ctx_lookup_len", "zpo_ctx_pi_pos", "zpo_ctx_wiring_len", "zpo_ctx_num_vars"]:
            getattr(L, name).restype = c64
            getattr(L, name).argtypes = [vp]
        L.zpo_ctx_logn.argtypes = [vp]
        for name in ["zpo_ctx_pi", "zpo_ctx_q_lookup", "zpo_ctx_linear_evaluations", "zpo_ctx_v_h_coset_8n", "zpo_ctx_srs",
                     "zpo_ctx_tau"]:
            getattr(L, name).restype = u64p
            getattr(L, name).argtypes = [vp]
        for name in ["zpo_ctx_wire", "zpo_ctx_selector_evals", "zpo_ctx_pk_coeffs", "zpo_ctx_pk_evals", "zpo_ctx_table"]:
            getattr(L, name).restype = u64p
            getattr(L, name).argtypes = [vp, ci]
        L.zpo_ctx_check_satisfied.argtypes = [vp]
        L.zpo_ctx_prove.restype = ctypes.c_double
        L.zpo_ctx_prove.argtypes = [vp, u64p, u64p]
        L.zpo_ctx_verify.argtypes = [vp, u64p, ctypes.POINTER(ci)]
        L.zpo_ctx_set_vk.argtypes = [vp, u64p]
        L.zpo_ctx_commit_with_tau.argtypes = [vp, cs, u64p, u64p]
        L.zpo_fr_op.argtypes = [ci, cs, u64p, u64p, u64p]
        L.zpo_fq_op.argtypes = [ci, cs, u64p, u64p, u64p]
        L.zpo_random_fr.argtypes = [c64, cs, u64p]
        L.zpo_ntt.argtypes = [ci, ci, u64p]
        L.zpo_poly_eval.argtypes = [cs, u64p, u64p, u64p]
        L.zpo_msm.argtypes = [cs, u64p, u64p, u64p]
        L.zpo_srs.argtypes = [c64, cs, u64p, u64p]
        L.zpo_g1_mul.argtypes = [u64p, u64p, u64p]
        L.zpo_g1_on_curve.argtypes = [u64p]
        L.zpo_g1_serialize.argtypes = [u64p, ctypes.c_void_p]
        L.zpo_transcript_kat.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, cs, ctypes.c_char_p,
                                         ctypes.c_void_p, cs]
        L.zpo_transcript_script.argtypes = [ctypes.c_char_p, ctypes.c_char_p, cs, ctypes.c_void_p]
        L.zpo_combine_split.argtypes = [cs, u64p, u64p, u64p, u64p]
        L.zpo_time_ntt.restype = ctypes.c_double
        L.zpo_time_ntt.argtypes = [ci, ci, ci, c64]
        L.zpo_time_msm.restype = ctypes.c_double
        L.zpo_time_msm.argtypes = [cs, u64p, u64p, u64p]
        L.zpo_fr_constants.argtypes = [u64p, u64p, u64p, u64p, u64p, u64p]
        L.zpo_fq_constants.argtypes = [u64p, u64p, u64p, u64p]
        L.zpo_fr_root_of_unity.argtypes = [ci, u64p, u64p, u64p]
        L.zpo_jubjub.argtypes = [u64p, u64p]
        L.zpo_g1_generator.argtypes = [u64p]

    # ---- field / poly helpers on numpy arrays of shape [n, 4] (Fr) or [n, 6] (Fq) ----
    def fr_op(self, op, a, b=None):
        out = np.zeros_like(a)
        self.lib.zpo_fr_op(op, a.shape[0], _p(a), _p(b), _p(out))
        return out

    def fq_op(self, op, a, b=None):
        out = np.zeros_like(a)
        self.lib.zpo_fq_op(op, a.shape[0], _p(a), _p(b), _p(out))
        return out

    def random_fr(self, seed, n):
        out = np.zeros((n, 4), dtype=np.uint64)
        self.lib.zpo_random_fr(seed, n, _p(out))
        return out

    def ntt(self, kind, data):
        out = data.copy()
        self.lib.zpo_ntt(kind, int(np.log2(data.shape[0])), _p(out))
        return out

    def poly_eval(self, coeffs, point):
        out = np.zeros(4, dtype=np.uint64)
        self.lib.zpo_poly_eval(coeffs.shape[0], _p(coeffs), _p(point), _p(out))
        return out

    def msm(self, points, scalars):
        out = np.zeros(12, dtype=np.uint64)
        self.lib.zpo_msm(scalars.shape[0], _p(points), _p(scalars), _p(out))
        return out

    def srs(self, tau_seed, n):
        pts = np.zeros((n, 12), dtype=np.uint64)
        tau = np.zeros(4, dtype=np.uint64)
        self.lib.zpo_srs(tau_seed, n, _p(pts), _p(tau))
        return pts, tau

    def g1_mul(self, point, scalar):
        out = np.zeros(12, dtype=np.uint64)
        self.lib.zpo_g1_mul(_p(point), _p(scalar), _p(out))
        return out

    def combine_split(self, t, f):
        h1, h2 = np.zeros_like(t), np.zeros_like(t)
        ok = self.lib.zpo_combine_split(t.shape[0], _p(t), _p(f), _p(h1), _p(h2))
        return ok, h1, h2

    def transcript_script(self, proto, ops):
        """ops: list of ('append', label, bytes) / ('challenge', label, nbytes) -> concatenated challenge bytes."""
        script, total = encode_script(ops)
        out = ctypes.create_string_buffer(max(total, 1))
        self.lib.zpo_transcript_script(proto, script, len(script), out)
        return out.raw[:total]


def encode_script(ops):
    import struct
    script = b""
    total = 0
    for op in ops:
        if op[0] == "append":
            script += b"\x00" + struct.pack("<I", len(op[1])) + op[1] + struct.pack("<I", len(op[2])) + op[2]
        else:
            script += b"\x01" + struct.pack("<I", len(op[1])) + op[1] + struct.pack("<I", op[2])
            total += op[2]
    return script, total


class OracleCircuit:
    """A synthetic Merkle-tree circuit + prover key + SRS held by the oracle, exposed as numpy views."""

    # kind: 0 Poseidon-Merkle tree of `height`; 1 all TurboPLONK widgets (range, logic, fixed-base, curve addition, q_m,
    # Poseidon rounds, optional lookups); 2 no constants (q_c == 0); 3 no arithmetic gates (q_arith == 0, no public input)
    def __init__(self, oracle, height, witness_seed=42, tau_seed=7, n_lookup=0, with_pk=True, with_srs=True, kind=0):
        self.o = oracle
        L = oracle.lib
        self.h = L.zpo_ctx_new_kind(kind, height, witness_seed, tau_seed, n_lookup, 1 if with_pk else 0, 1 if with_srs else 0)
        self.cs_n = L.zpo_ctx_n(self.h)
        self.log_n = L.zpo_ctx_logn(self.h)
        self.n = 1 << self.log_n
        self.lookup_len = L.zpo_ctx_lookup_len(self.h)
        self.pi_pos = L.zpo_ctx_pi_pos(self.h)
        self.with_pk, self.with_srs = with_pk, with_srs

    def _view(self, ptr, rows, cols=4):
        return np.ctypeslib.as_array(ptr, shape=(rows, cols)).copy()

    def wires(self):
        return [self._view(self.o.lib.zpo_ctx_wire(self.h, k), self.cs_n) for k in range(4)]

    def q_lookup(self):
        return self._view(self.o.lib.zpo_ctx_q_lookup(self.h), self.cs_n)

    def pi_canonical(self):
        return self._view(self.o.lib.zpo_ctx_pi(self.h), 1).reshape(4)

    def selector_evals(self):
        return [self._view(self.o.lib.zpo_ctx_selector_evals(self.h, s), self.n) for s in range(19)]

    def tables(self):
        return [self._view(self.o.lib.zpo_ctx_table(self.h, c), self.n) for c in range(4)]

    def pk_coeffs(self):
        return [self._view(self.o.lib.zpo_ctx_pk_coeffs(self.h, s), self.n) for s in range(19)]

    def pk_evals(self):
        return [self._view(self.o.lib.zpo_ctx_pk_evals(self.h, s), 8 * self.n) for s in range(19)]

    def linear_evaluations(self):
        return self._view(self.o.lib.zpo_ctx_linear_evaluations(self.h), 8 * self.n)

    def v_h_coset_8n(self):
        return self._view(self.o.lib.zpo_ctx_v_h_coset_8n(self.h), 8 * self.n)

    def srs(self):
        return self._view(self.o.lib.zpo_ctx_srs(self.h), self.n, 12)

    def wiring(self):
        """(vars, cells, n_vars): the composer's wire map in insertion order; cell = (gate << 2) | wire."""
        m = self.o.lib.zpo_ctx_wiring_len(self.h)
        v, c = np.zeros(m, dtype=np.uint32), np.zeros(m, dtype=np.uint32)
        self.o.lib.zpo_ctx_wiring(self.h, v.ctypes.data, c.ctypes.data)
        return v, c, int(self.o.lib.zpo_ctx_num_vars(self.h))

    def tau(self):
        return self._view(self.o.lib.zpo_ctx_tau(self.h), 1).reshape(4)

    def satisfied(self):
        return self.o.lib.zpo_ctx_check_satisfied(self.h) == 1

    def prove(self):
        proof = np.zeros(332, dtype=np.uint64)
        secs = self.o.lib.zpo_ctx_prove(self.h, _p(proof), None)
        return proof, secs

    def verify(self, proof_words):
        d = ctypes.c_int()
        pw = np.ascontiguousarray(proof_words, dtype=np.uint64)
        ok = self.o.lib.zpo_ctx_verify(self.h, _p(pw), ctypes.byref(d))
        return ok == 1, d.value

    def set_vk(self, comms23):
        self.o.lib.zpo_ctx_set_vk(self.h, _p(np.ascontiguousarray(comms23, dtype=np.uint64)))

    def commit_with_tau(self, coeffs):
        out = np.zeros(12, dtype=np.uint64)
        self.o.lib.zpo_ctx_commit_with_tau(self.h, coeffs.shape[0], _p(coeffs), _p(out))
        return out

    def close(self):
        if self.h:
            self.o.lib.zpo_ctx_free(self.h)
            self.h = None


_oracle = None


def load():
    global _oracle
    if _oracle is None:
        subprocess.run([os.path.join(ORACLE_DIR, "build.sh")], check=True, stdout=subprocess.DEVNULL)
        _oracle = Oracle(ctypes.CDLL(os.path.join(ORACLE_DIR, "liboracle.so")))
    return _oracle


def ref_lib(name):
    """oracle/_ref/<name> (the reference's own sources compiled here) or None when absent."""
    path = os.path.join(ORACLE_DIR, "_ref", name)
    return ctypes.CDLL(path) if os.path.exists(path) else None
