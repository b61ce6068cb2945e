"""ctypes view of the CPU oracle (oracle/oracle.cc -> oracle/_build/liboracle.so).

TEST INFRASTRUCTURE: the checker, never the thing measured or shipped.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; nothing under zkir_b200/ does."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_oracle():
    import subprocess
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "oracle.cc")
    if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return so


class Oracle:
    """ctypes view of oracle/_build/liboracle.so -- the checker, never the thing under test."""

    def __init__(self):
        self.l = C.CDLL(_build_oracle())
        self.l.oracle_proof_words.restype = C.c_uint64

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def params(self, cfg):
        return np.array([cfg.log_blowup, cfg.num_queries, cfg.pow_bits, 72, 4], dtype=np.uint32)

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

    def quotient(self, cfg, lde, log_n, pv, alpha):
        M = lde.shape[1]
        out = np.empty((4, M), dtype=np.uint32)
        pr = self.params(cfg)
        pv = np.ascontiguousarray(pv, dtype=np.uint32)
        al = np.ascontiguousarray(alpha, dtype=np.uint32)
        self.l.oracle_quotient(self._p(pr), log_n, self._p(np.ascontiguousarray(lde)), self._p(pv), self._p(al), self._p(out))
        return out

    def fri_fold(self, layer, shift, beta):
        a = np.ascontiguousarray(layer, dtype=np.uint32)
        n = a.shape[0]
        out = np.empty((n // 2, 4), dtype=np.uint32)
        b = np.ascontiguousarray(beta, dtype=np.uint32)
        self.l.oracle_fri_fold(self._p(a), self._p(out), int(n).bit_length() - 1, C.c_uint32(shift), self._p(b))
        return out

    def check_trace(self, cols, pv):
        bad = C.c_uint64()
        cols = np.ascontiguousarray(cols)
        k = self.l.oracle_check_trace(self._p(cols), int(cols.shape[1]).bit_length() - 1, self._p(np.ascontiguousarray(pv, dtype=np.uint32)), C.byref(bad))
        return k, bad.value

    def prove(self, cfg, cols, pv):
        pr = self.params(cfg)
        cols = np.ascontiguousarray(cols)
        log_n = int(cols.shape[1]).bit_length() - 1
        nw = self.l.oracle_proof_words(self._p(pr), log_n)
        proof = np.zeros(nw, dtype=np.uint32)
        rc = self.l.oracle_prove(self._p(pr), self._p(cols), log_n, self._p(np.ascontiguousarray(pv, dtype=np.uint32)), self._p(proof))
        assert rc == 0, f"oracle_prove failed rc={rc}"
        return proof.tobytes()
