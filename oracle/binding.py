"""ctypes view of the CPU oracle (oracle/oracle.cc -> oracle/_build/liboracle.so).

TEST INFRASTRUCTURE: the checker, never the thing measured or shipped.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; nothing under zkir_b200/ does."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_oracle(full=False):
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "liboracle_full.so" if full else "liboracle.so")
    src = os.path.join(ROOT, "oracle", "oracle.cc")
    if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return so


def _profile_numbers():
    """the #defines of oracle/air_profiles_generated.h (tools/gen_air.py)"""
    out = {}
    for line in open(os.path.join(ROOT, "oracle", "air_profiles_generated.h")):
        f = line.split()
        if len(f) == 3 and f[0] == "#define":
            out[f[1]] = int(f[2])
    return out


class Oracle:
    """ctypes view of oracle/_build/liboracle.so -- the checker, never the thing under test."""

    def __init__(self, width=88):
        """`width` selects the AIR profile (docs/PROVER_SPEC.md section 3.7): 88 = core (liboracle.so), FULL_WIDTH = full (liboracle_full.so)"""
        full = width == self.FULL_WIDTH
        assert full or width == 88
        self.l = C.CDLL(_build_oracle(full))
        self.l.oracle_proof_words.restype = C.c_uint64
        if full:
            self.WIDTH, self.AUX_WIDTH, self.PUB_WIDTH = self.FULL_WIDTH, self.FULL_AUX_WIDTH, self.FULL_PUB_WIDTH

    @classmethod
    def for_columns(cls, cols):
        return cls(width=int(cols.shape[0]))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    WIDTH, AUX_WIDTH, PUB_WIDTH, NUM_PUBLIC = 88, 16, 4, 5   # AIR v2, core profile (oracle/air_generated.h)
    FULL_WIDTH, FULL_AUX_WIDTH, FULL_PUB_WIDTH = (_profile_numbers()[f"ZKIR_PROFILE_FULL_{k}"] for k in ("WIDTH", "AUX", "PUB"))   # full profile
    LOOKUP_TEST = np.array([3, 1, 4, 1, 5, 9, 2, 6], dtype=np.uint32)   # fixed lookup challenges z, theta for row-domain checks

    def params(self, cfg):
        return np.array([cfg.log_blowup, cfg.num_queries, cfg.pow_bits, self.WIDTH, self.NUM_PUBLIC], dtype=np.uint32)

    @staticmethod
    def _code(program):
        """`program`: Program, code words, or an ExecutionResult (anything with .program)"""
        program = getattr(program, "program", program)
        return np.ascontiguousarray(getattr(program, "code", program), dtype=np.uint32)

    @staticmethod
    def _io(program, io):
        """public I/O transcript uint32 [n, 4]: explicit, or the ExecutionResult's own, or empty"""
        if io is None:
            io = getattr(program, "io", None) if hasattr(program, "program") else None
        return np.ascontiguousarray(io if io is not None else np.zeros((0, 4)), dtype=np.uint32).reshape(-1, 4)

    def ntt(self, cols, inverse=False):
        a = np.ascontiguousarray(cols, dtype=np.uint32).copy()
        n_cols, n = a.shape
        self.l.oracle_ntt_batch(self._p(a), n_cols, int(n).bit_length() - 1, int(inverse))
        return a

    def lde(self, cols, log_blowup):
        a = np.ascontiguousarray(cols, dtype=np.uint32)
        n_cols, n = a.shape
        out = np.empty((n_cols, n << log_blowup), dtype=np.uint32)
        self.l.oracle_lde_batch(self._p(a), self._p(out), n_cols, int(n).bit_length() - 1, log_blowup)
        return out

    def poseidon2(self, states):
        a = np.ascontiguousarray(states, dtype=np.uint32).copy()
        self.l.oracle_poseidon2(self._p(a), C.c_uint64(a.shape[0]))
        return a

    def hash_tree(self, words):
        a = np.ascontiguousarray(words, dtype=np.uint32)
        d = np.empty(8, dtype=np.uint32)
        self.l.oracle_hash_tree(self._p(a), C.c_uint64(a.shape[0]), self._p(d))
        return d

    def merkle_commit(self, mat):
        a = np.ascontiguousarray(mat, dtype=np.uint32)
        n_cols, rows = a.shape
        tree = np.empty((2 * rows - 1, 8), dtype=np.uint32)
        root = np.empty(8, dtype=np.uint32)
        self.l.oracle_merkle_commit(self._p(a), n_cols, int(rows).bit_length() - 1, self._p(tree), self._p(root))
        return tree, root

    def quotient(self, cfg, lde, publde, log_n, pv, lookup, alpha, io=None):
        """lde [WIDTH + 16][M] (main then aux columns), publde [4][M], natural order -> quotient values [4][M]"""
        M = lde.shape[1]
        out = np.empty((4, M), dtype=np.uint32)
        pr = self.params(cfg)
        pv = np.ascontiguousarray(pv, dtype=np.uint32)
        al = np.ascontiguousarray(alpha, dtype=np.uint32)
        lk = np.ascontiguousarray(lookup, dtype=np.uint32)
        ev = self._io(None, io)
        self.l.oracle_quotient(self._p(pr), log_n, self._p(np.ascontiguousarray(lde)), self._p(np.ascontiguousarray(publde)), self._p(pv), self._p(lk),
                               self._p(ev), C.c_uint64(ev.shape[0]), self._p(al), self._p(out))
        return out

    def public_columns(self, log_n, program):
        code = self._code(program)
        pub = np.empty((self.PUB_WIDTH, 1 << log_n), dtype=np.uint32)
        self.l.oracle_public_columns(log_n, self._p(code), C.c_uint64(len(code)), self._p(pub))
        return pub

    def aux_columns(self, cols, program, lookup=None, io=None):
        """-> (aux [16][N], balanced?) for the lookup challenges z, theta"""
        cols = np.ascontiguousarray(cols)
        code = self._code(program)
        ev = self._io(program, io)
        lk = np.ascontiguousarray(self.LOOKUP_TEST if lookup is None else lookup, dtype=np.uint32)
        aux = np.empty((self.AUX_WIDTH, cols.shape[1]), dtype=np.uint32)
        ok = self.l.oracle_aux_columns(int(cols.shape[1]).bit_length() - 1, self._p(cols), self._p(code), C.c_uint64(len(code)), self._p(ev), C.c_uint64(ev.shape[0]),
                                       self._p(lk), self._p(aux))
        return aux, bool(ok)

    def program_digest(self, program):
        code = self._code(program)
        d = np.empty(8, dtype=np.uint32)
        self.l.oracle_program_digest(self._p(code), C.c_uint64(len(code)), self._p(d))
        return d

    def fri_fold(self, layer, shift, beta):
        a = np.ascontiguousarray(layer, dtype=np.uint32)
        n = a.shape[0]
        out = np.empty((n // 2, 4), dtype=np.uint32)
        b = np.ascontiguousarray(beta, dtype=np.uint32)
        self.l.oracle_fri_fold(self._p(a), self._p(out), int(n).bit_length() - 1, C.c_uint32(shift), self._p(b))
        return out

    def check_trace(self, cols, pv, program, lookup=None, io=None):
        """every AIR constraint (LogUp included, aux columns built for `lookup`) on the unextended rows:
        (-1, _) all hold; (-2, _) lookups unbalanced; (k, row) first failing constraint"""
        bad = C.c_uint64()
        cols = np.ascontiguousarray(cols)
        code = self._code(program)
        ev = self._io(program, io)
        lk = np.ascontiguousarray(self.LOOKUP_TEST if lookup is None else lookup, dtype=np.uint32)
        k = self.l.oracle_check_trace(self._p(cols), int(cols.shape[1]).bit_length() - 1, self._p(np.ascontiguousarray(pv, dtype=np.uint32)),
                                      self._p(code), C.c_uint64(len(code)), self._p(ev), C.c_uint64(ev.shape[0]), self._p(lk), C.byref(bad))
        return k, bad.value

    def prove(self, cfg, cols, pv, program, io=None):
        pr = self.params(cfg)
        cols = np.ascontiguousarray(cols)
        code = self._code(program)
        ev = self._io(program, io)
        log_n = int(cols.shape[1]).bit_length() - 1
        nw = self.l.oracle_proof_words(self._p(pr), log_n)
        proof = np.zeros(nw, dtype=np.uint32)
        rc = self.l.oracle_prove(self._p(pr), self._p(cols), log_n, self._p(np.ascontiguousarray(pv, dtype=np.uint32)), self._p(code), C.c_uint64(len(code)),
                                 self._p(ev), C.c_uint64(ev.shape[0]), self._p(proof))
        assert rc == 0, f"oracle_prove failed rc={rc}"
        return proof.tobytes()
