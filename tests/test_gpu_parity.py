"""GPU parity tests proper: every call goes through the C ABI (ctypes) and is compared bit-for-bit with the CPU
oracle on the same seeded inputs.  Integer field arithmetic => the bar is bit-exact everywhere."""
import numpy as np
import pytest

import zkir_b200
from conftest import P, fib_trace

pytestmark = pytest.mark.gpu


def rand_field(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint64).astype(np.uint32)


@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8, 10, 11, 12, 13, 16, 20])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_matches_oracle(gpu_ctx, oracle, log_n, inverse):
    rng = np.random.default_rng(1000 + log_n)
    n_cols = 3 if log_n >= 16 else 37   # ragged column count on purpose
    a = rand_field(rng, (n_cols, 1 << log_n))
    d = gpu_ctx.to_device(a)
    gpu_ctx.ntt(d, n_cols, log_n, inverse=inverse)
    got = gpu_ctx.to_host(d, a.shape)
    gpu_ctx.free(d)
    assert np.array_equal(got, oracle.ntt(a, inverse))


@pytest.mark.parametrize("log_n", [4, 9, 12, 17, 21, 23])
def test_ntt_roundtrip(gpu_ctx, log_n):
    rng = np.random.default_rng(7 + log_n)
    a = rand_field(rng, (2, 1 << log_n))
    d = gpu_ctx.to_device(a)
    gpu_ctx.ntt(d, 2, log_n, inverse=False)
    gpu_ctx.ntt(d, 2, log_n, inverse=True)
    got = gpu_ctx.to_host(d, a.shape)
    gpu_ctx.free(d)
    assert np.array_equal(got, a)


@pytest.mark.parametrize("log_n", [8, 9, 14, 19])
def test_coset_ntt_matches_oracle(gpu_ctx, oracle, log_n):
    """forward NTT on the coset 31*H: scale by 31^j then transform (docs/PROVER_SPEC.md section 4 step 1)."""
    rng = np.random.default_rng(300 + log_n)
    a = rand_field(rng, (3, 1 << log_n))
    d = gpu_ctx.to_device(a)
    gpu_ctx.ntt(d, 3, log_n, inverse=False, coset_shift=31)
    got = gpu_ctx.to_host(d, a.shape)
    gpu_ctx.free(d)
    pw = np.ones(1 << log_n, dtype=object)
    for j in range(1, 1 << log_n):
        pw[j] = pw[j - 1] * 31 % P
    scaled = ((a.astype(object) * pw) % P).astype(np.uint32)
    assert np.array_equal(got, oracle.ntt(scaled, False))


@pytest.mark.parametrize("log_n,log_b", [(2, 1), (6, 1), (8, 1), (9, 3), (10, 1), (10, 2), (11, 4), (12, 1), (15, 2), (18, 1), (21, 1)])
def test_lde_matches_oracle(gpu_ctx, oracle, log_n, log_b):
    rng = np.random.default_rng(50 + log_n)
    n_cols = 5 if log_n < 20 else 2
    a = rand_field(rng, (n_cols, 1 << log_n))
    d_in = gpu_ctx.to_device(a)
    d_out = gpu_ctx.alloc(n_cols * (4 << (log_n + log_b)))
    gpu_ctx.lde(d_in, d_out, n_cols, log_n, log_b)
    got = gpu_ctx.to_host(d_out, (n_cols, 1 << (log_n + log_b)))
    gpu_ctx.free(d_in); gpu_ctx.free(d_out)
    assert np.array_equal(got, oracle.lde(a, log_b))


def test_poseidon2_matches_oracle(gpu_ctx, oracle):
    rng = np.random.default_rng(3)
    st = rand_field(rng, (5000, 16))
    st[0] = 0
    st[1] = P - 1
    d = gpu_ctx.to_device(st)
    gpu_ctx.poseidon2_permute(d, st.shape[0])
    got = gpu_ctx.to_host(d, st.shape)
    gpu_ctx.free(d)
    assert np.array_equal(got, oracle.poseidon2(st))


@pytest.mark.parametrize("n_cols,log_rows", [(1, 1), (8, 4), (13, 7), (112, 12), (8, 15)])
def test_merkle_commit_matches_oracle(gpu_ctx, oracle, n_cols, log_rows):
    rng = np.random.default_rng(11 * n_cols + log_rows)
    m = rand_field(rng, (n_cols, 1 << log_rows))
    d = gpu_ctx.to_device(m)
    d_tree = gpu_ctx.alloc(((2 << log_rows) - 1) * 32)
    root = gpu_ctx.merkle_commit(d, n_cols, log_rows, d_tree)
    tree = gpu_ctx.to_host(d_tree, ((2 << log_rows) - 1, 8))
    gpu_ctx.free(d); gpu_ctx.free(d_tree)
    otree, oroot = oracle.merkle_commit(m)
    assert np.array_equal(root, oroot)
    assert np.array_equal(tree, otree)


@pytest.mark.parametrize("log_n", [1, 2, 7, 14])
def test_fri_fold_matches_oracle(gpu_ctx, oracle, log_n):
    rng = np.random.default_rng(90 + log_n)
    layer = rand_field(rng, (1 << log_n, 4))
    beta = rand_field(rng, 4)
    shift = 31
    d_in = gpu_ctx.to_device(layer)
    d_out = gpu_ctx.alloc(16 << (log_n - 1) if log_n > 1 else 16)
    gpu_ctx.fri_fold(d_in, d_out, log_n, shift, beta)
    got = gpu_ctx.to_host(d_out, (1 << (log_n - 1), 4))
    gpu_ctx.free(d_in); gpu_ctx.free(d_out)
    assert np.array_equal(got, oracle.fri_fold(layer, shift, beta))


@pytest.mark.parametrize("n,log_b", [(10, 1), (300, 1), (30, 2)])
def test_quotient_matches_oracle(gpu_ctx, oracle, n, log_b):
    """all 169 constraints of the AIR v2 (main + LogUp aux + public columns) on the LDE coset, folded with alpha, divided by Z_H"""
    res, cols, pv = fib_trace(n)
    cfg = zkir_b200.ProverConfig(log_blowup=log_b, num_queries=4, pow_bits=1)
    log_n = int(cols.shape[1]).bit_length() - 1
    lookup = np.array([11, 12, 13, 14, 21, 22, 23, 24], dtype=np.uint32)
    aux, balanced = oracle.aux_columns(cols, res, lookup)
    assert balanced
    lde = oracle.lde(np.concatenate([cols, aux]), log_b)
    publde = oracle.lde(oracle.public_columns(log_n, res), log_b)
    alpha = np.array([5, 6, 7, 8], dtype=np.uint32)
    d_lde, d_pub = gpu_ctx.to_device(lde), gpu_ctx.to_device(publde)
    d_q = gpu_ctx.alloc(16 << (log_n + log_b))
    gpu_ctx.set_io(res.io)
    gpu_ctx.quotient(cfg, d_lde, d_pub, log_n, pv, lookup, alpha, d_q)
    got = gpu_ctx.to_host(d_q, (4, 1 << (log_n + log_b)))
    gpu_ctx.free(d_lde); gpu_ctx.free(d_pub); gpu_ctx.free(d_q)
    assert np.array_equal(got, oracle.quotient(cfg, lde, publde, log_n, pv, lookup, alpha, io=res.io))


@pytest.mark.parametrize("case", ["fib30", "fib300", "family"])
def test_aux_columns_match_oracle(gpu_ctx, oracle, case):
    """LogUp aux columns (three helper sums + running sum, ext4 as 4 base columns each) for fixed lookup challenges"""
    if case == "family":
        from test_oracle_cpu import FAMILY_SRC
        prog = zkir_b200.assemble(FAMILY_SRC.format(jalr_imm=4 * 29 + 1))
        res = zkir_b200.VM(prog, [(5 << 20) + 3, 9], zkir_b200.VMConfig(enable_execution_trace=True)).run()
        cols, pv = res.pack(12)
    else:
        res, cols, pv = fib_trace(int(case[3:]))
    log_n = int(cols.shape[1]).bit_length() - 1
    lookup = np.array([101, 2, 3, 4, 5, 6, 7, 8], dtype=np.uint32)
    want, balanced = oracle.aux_columns(cols, res, lookup)
    assert balanced
    gpu_ctx.set_program(res)
    d_t = gpu_ctx.to_device(cols)
    d_a = gpu_ctx.alloc(want.nbytes)
    gpu_ctx.aux_columns(d_t, log_n, lookup, d_a)
    got = gpu_ctx.to_host(d_a, want.shape)
    gpu_ctx.free(d_t); gpu_ctx.free(d_a)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first mismatch aux column {bad[0][0]} row {bad[0][1]}"
    # a trace whose lookups cannot balance (a chunk outside the table) is reported, not proven
    L = zkir_b200.air_layout.INDEX
    t2 = cols.copy(); t2[L["ch0"], 4] += 1024
    d_t = gpu_ctx.to_device(t2)
    d_a = gpu_ctx.alloc(want.nbytes)
    with pytest.raises(zkir_b200.RuntimeError) as ei:
        gpu_ctx.aux_columns(d_t, log_n, lookup, d_a)
    gpu_ctx.free(d_t); gpu_ctx.free(d_a)
    assert ei.value.code == -6 and "lookup" in str(ei.value)


@pytest.mark.parametrize("n,log_b,nq,pow_bits", [(3, 1, 3, 0), (30, 1, 100, 16), (30, 2, 20, 4), (205, 1, 100, 16), (300, 1, 30, 10), (1000, 1, 30, 10),
                                                    (205, 2, 12, 5), (205, 3, 8, 3), (1000, 2, 10, 6), (3000, 4, 6, 2)])
def test_proof_bytes_match_oracle_and_verify(gpu_ctx, oracle, n, log_b, nq, pow_bits):
    """BASELINE config 1 (fib n=30 / n=205): whole proof bytes GPU == oracle, and the verifier accepts."""
    res, cols, pv = fib_trace(n)
    cfg = zkir_b200.ProverConfig(log_blowup=log_b, num_queries=nq, pow_bits=pow_bits)
    got = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    want = oracle.prove(cfg, cols, pv, res)
    assert len(got) == len(want)
    if got != want:
        g, w = np.frombuffer(got, dtype=np.uint32), np.frombuffer(want, dtype=np.uint32)
        first = int(np.nonzero(g != w)[0][0])
        pytest.fail(f"proof words differ first at word {first}: gpu={g[first]} oracle={w[first]}")
    ok, why = zkir_b200.verify(got, cfg, pv, res)
    assert ok, why


ROW_PROGRAMS = {
    "fib30": ("fib", 30, []),
    "fib_input": ("fib_input", None, [77]),
    "mixed": ("src", """
addi r10, r0, 1
ecall
add r1, r10, r0
addi r10, r0, 1
ecall
sub r2, r1, r10
sub r3, r10, r1
beq r2, r3, 8
addi r4, r0, -5
jal r5, 8
addi r6, r0, 9
bne r1, r10, 8
addi r7, r0, 1
add r11, r2, r0
addi r10, r0, 2
ecall
addi r11, r0, 3
addi r10, r0, 0
ecall
""", [123456789012, 7]),
}


from test_oracle_cpu import FAMILY_SRC  # noqa: E402
ROW_PROGRAMS["family_lt"] = ("src", FAMILY_SRC.format(jalr_imm=4 * 29 + 1), [3, (5 << 20) + 3])
ROW_PROGRAMS["family_eq"] = ("src", FAMILY_SRC.format(jalr_imm=4 * 29), [(1 << 40) - 1, (1 << 40) - 1])
ROW_PROGRAMS["ebreak"] = ("src", "addi r1, r0, 5\nebreak\n", [])
ROW_PROGRAMS["read_one"] = ("src", "addi r10, r0, 1\necall\nadd r3, r10, r10\naddi r10, r0, 1\necall\nadd r10, r0, r0\necall\n", [1, 1])


def _rows_case(name):
    from conftest import fib_program, fib_program_input
    kind, arg, inputs = ROW_PROGRAMS[name]
    prog = fib_program(arg) if kind == "fib" else (fib_program_input() if kind == "fib_input" else zkir_b200.assemble(arg))
    return zkir_b200.VM(prog, inputs, zkir_b200.VMConfig(max_cycles=1 << 20, enable_execution_trace=True)).run()


@pytest.mark.parametrize("name", sorted(ROW_PROGRAMS))
def test_expand_rows_matches_host_packer(gpu_ctx, name):
    """device converter (trace_expand.cu) == host converter (pack.cc), bit for bit, including padding rows"""
    res = _rows_case(name)
    for log_n in (res.min_log_n(), res.min_log_n() + 2):
        cols, pv = res.pack(log_n)
        d = gpu_ctx.alloc(cols.nbytes)
        gpu_ctx.expand_rows(res.rows(), log_n, d)
        got = gpu_ctx.to_host(d, cols.shape)
        gpu_ctx.free(d)
        bad = np.argwhere(got != cols)
        assert bad.size == 0, f"first mismatch column {zkir_b200.air_layout.COLUMNS[bad[0][0]]} row {bad[0][1]}: gpu={got[tuple(bad[0])]} host={cols[tuple(bad[0])]}"


@pytest.mark.parametrize("name", sorted(ROW_PROGRAMS))
def test_expand_writelog_matches_host_packer(gpu_ctx, name):
    """register write log (16 B/row) -> last-writer scan -> converter == host converter on the full rows"""
    res = _rows_case(name)
    for log_n in (res.min_log_n(), res.min_log_n() + 1, 12):
        cols, pv = res.pack(log_n)
        d = gpu_ctx.alloc(cols.nbytes)
        gpu_ctx.expand_writelog(res.writelog(), log_n, d)
        got = gpu_ctx.to_host(d, cols.shape)
        gpu_ctx.free(d)
        bad = np.argwhere(got != cols)
        assert bad.size == 0, f"first mismatch column {zkir_b200.air_layout.COLUMNS[bad[0][0]]} row {bad[0][1]}: gpu={got[tuple(bad[0])]} host={cols[tuple(bad[0])]}"


def test_expand_writelog_long_trace(gpu_ctx):
    """many chunks: exercises the cross-chunk scan (2^16 rows = 256 chunks of 256 rows)"""
    _, cols, pv = fib_trace(n_input=13000, log_n=16)
    from conftest import fib_program_input
    res = zkir_b200.VM(fib_program_input(), [13000], zkir_b200.VMConfig(max_cycles=1 << 20, enable_execution_trace=True)).run()
    d = gpu_ctx.alloc(cols.nbytes)
    gpu_ctx.expand_writelog(res.writelog(), 16, d)
    got = gpu_ctx.to_host(d, cols.shape)
    gpu_ctx.free(d)
    assert np.array_equal(got, cols)


def test_expand_rows_rejects_unconstrained_opcode(gpu_ctx):
    prog = zkir_b200.assemble("addi r1, r0, 3\nmul r2, r1, r1\nadd r10, r0, r0\necall\n")
    res = zkir_b200.VM(prog, [], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    d = gpu_ctx.alloc(zkir_b200.air_layout.WIDTH * 4 << 10)
    with pytest.raises(zkir_b200.RuntimeError) as ei:
        gpu_ctx.expand_rows(res.rows(), 10, d)
    gpu_ctx.free(d)
    assert ei.value.code == -6 and "row 1" in str(ei.value)


@pytest.mark.parametrize("name", sorted(ROW_PROGRAMS))
def test_prove_rows_equals_prove_columns(gpu_ctx, oracle, name):
    res = _rows_case(name)
    cfg = zkir_b200.ProverConfig(num_queries=12, pow_bits=6)
    cols, pv = res.pack()
    from_cols = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    from_rows, pv2 = gpu_ctx.prove_rows(res.rows(), cfg)
    from_wl, pv3 = gpu_ctx.prove_writelog(res.writelog(), cfg)
    assert np.array_equal(pv, pv2) and np.array_equal(pv, pv3)
    assert from_rows == from_cols == from_wl == oracle.prove(cfg, cols, pv, res)
    ok, why = zkir_b200.verify(from_wl, cfg, pv, res)
    assert ok, why


def test_prove_api_end_to_end(gpu_ctx):
    from conftest import fib_program
    cfg = zkir_b200.ProverConfig(num_queries=20, pow_bits=8)
    proof = zkir_b200.prove(fib_program(30), [], cfg)
    assert proof.cycles == 146 and proof.log_n == 10
    ok, why = zkir_b200.verify(proof, cfg)
    assert ok, why
    bad = bytearray(proof.bytes_)
    bad[-5] ^= 1
    ok, _ = zkir_b200.verify(bytes(bad), cfg, proof.public_values)
    assert not ok


@pytest.mark.parametrize("n_input,log_n", [(13000, 16), (209715, 20)])
def test_full_size_proof_bytes_match_oracle(gpu_ctx, oracle, n_input, log_n):
    """BASELINE config 2 itself (fibonacci n=209715: 1_048_573 cycles, 2^20 rows, default parameters) and its 2^16-row
    sibling: the WHOLE proof, byte for byte, GPU == CPU oracle (OpenMP; about 10 s at 2^20), through every entry point the
    benchmark times (resident columns, raw rows, register write log)."""
    res, cols, pv = fib_trace(n_input=n_input)
    assert cols.shape[1] == 1 << log_n and res.cycles == 5 * n_input - 2
    cfg = zkir_b200.ProverConfig()
    want = oracle.prove(cfg, cols, pv, res)
    got = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    assert len(got) == len(want)
    if got != want:
        g, w = np.frombuffer(got, dtype=np.uint32), np.frombuffer(want, dtype=np.uint32)
        first = int(np.nonzero(g != w)[0][0])
        pytest.fail(f"2^{log_n}-row proof differs from the oracle first at word {first}: gpu={g[first]} oracle={w[first]}")
    from_wl, pv_wl = gpu_ctx.prove_writelog(res.writelog(), cfg, log_n)
    assert from_wl == want and list(pv_wl) == list(pv)
    ok, why = zkir_b200.verify(got, cfg, pv, res)
    assert ok, why
    if log_n == 20:   # Program -> Proof in one call (interpreter + overlapped upload + proof): the same bytes again
        from conftest import fib_program_input
        cfg2 = zkir_b200.ProverConfig(max_cycles=1 << 20)
        pb, pv2, cycles, ln = gpu_ctx.prove_program(fib_program_input(), [n_input], cfg2)
        assert (cycles, ln) == (res.cycles, 20) and pb == want and list(pv2) == list(pv)


def test_writelog_rejects_values_above_40_bits(gpu_ctx):
    """A caller that logs the reference's unmasked u64 register writes must get ZKIR_ERR_AIR from the write-log path exactly as
    from the full-row path -- never a proof of a truncated execution (bits 40..55 of a log word, bits 60..63, or a payload
    without a register index)."""
    res = _rows_case("mixed")
    cfg = zkir_b200.ProverConfig(num_queries=4, pow_bits=2)
    good = res.writelog()
    row = int(np.nonzero(good["wlog"] >> np.uint64(56))[0][3])       # some row that writes a register
    for bad_word in (good["wlog"][row] | np.uint64(1 << 41), good["wlog"][row] | np.uint64(1 << 61), np.uint64(5)):
        wl = dict(good, wlog=good["wlog"].copy())
        wl["wlog"][row] = bad_word
        with pytest.raises(zkir_b200.RuntimeError) as ei:
            gpu_ctx.prove_writelog(wl, cfg)
        assert ei.value.code == -6 and f"row {row}" in str(ei.value)
    rows = res.rows()
    regs = rows["regs"].copy()
    regs[row + 1, int(good["wlog"][row] >> np.uint64(56))] |= np.uint64(1 << 41)
    with pytest.raises(zkir_b200.RuntimeError) as ei:
        gpu_ctx.prove_rows(dict(rows, regs=regs), cfg)
    assert ei.value.code == -6
    assert gpu_ctx.prove_writelog(good, cfg)[0] == gpu_ctx.prove_rows(rows, cfg)[0]    # the context is still healthy


def test_large_trace_proves_and_verifies(gpu_ctx):
    """2^16-row trace: too slow for a byte comparison with the scalar oracle in CI time, so use the size-independent
    property: the independent CPU verifier accepts and rejects a flipped bit."""
    res, cols, pv = fib_trace(n_input=13000, log_n=16)
    cfg = zkir_b200.ProverConfig()
    pb = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    ok, why = zkir_b200.verify(pb, cfg, pv, res)
    assert ok, why


def test_2p22_row_trace_proves_and_verifies(gpu_ctx):
    """BASELINE config 3 size (2^22 rows; three NTT digits 8+7+7): raw rows in, device converter, proof accepted by the
    independent CPU verifier, and a flipped proof bit rejected.  (Byte comparison with the scalar oracle would take minutes.)"""
    from conftest import fib_program_input
    n = 838860                                   # 5n - 2 = 4_194_298 cycles <= 2^22
    res = zkir_b200.VM(fib_program_input(), [n], zkir_b200.VMConfig(max_cycles=1 << 23, enable_execution_trace=True)).run()
    assert res.cycles == 5 * n - 2 and res.min_log_n() == 22
    cfg = zkir_b200.ProverConfig(num_queries=40, pow_bits=12)
    pb, pv = gpu_ctx.prove_rows(res.rows(), cfg, 22)
    assert list(pv[:2]) == [0x1000, (5 * n - 2) % P]
    ok, why = zkir_b200.verify(pb, cfg, pv, res)
    assert ok, why
    bad = bytearray(pb)
    bad[len(bad) // 2] ^= 4
    ok, _ = zkir_b200.verify(bytes(bad), cfg, pv, res)
    assert not ok


from conftest import ADD_SRC  # noqa: E402


def test_prove_batch_matches_single_proofs(gpu_ctx, oracle):
    """BASELINE config 4 (a+b, 11 cycles -> 2^10 rows: the range table sets the minimum; many independent proofs): the concurrent batch path returns, for
    every i, exactly the bytes a single zkir_b200_prove gives, which equal the oracle's."""
    prog = zkir_b200.assemble(ADD_SRC)
    cfg = zkir_b200.ProverConfig(num_queries=16, pow_bits=4)
    traces = []
    for i in range(24):
        res = zkir_b200.VM(prog, [i, 2 * i + 1], zkir_b200.VMConfig(enable_execution_trace=True)).run()
        assert res.outputs == [3 * i + 1] and res.cycles == 11
        assert res.io.tolist() == [[1, 0, i, 0], [4, 0, 2 * i + 1, 0], [7, 1, 3 * i + 1, 0]]     # (cycle, kind, lo, hi) of the two READs and the WRITE
        traces.append(res.pack() + (res,))
    batch = gpu_ctx.prove_batch([c for c, _, _ in traces], [p for _, p, _ in traces], cfg, program=prog, io_list=[r.io for _, _, r in traces])
    assert len(batch) == 24
    for i, (cols, pv, res) in enumerate(traces):
        assert batch[i] == gpu_ctx.prove_columns(cols, pv, cfg, program=res)
        ok, why = zkir_b200.verify(batch[i], cfg, pv, res)
        assert ok, why
        ok, _ = zkir_b200.verify(batch[i], cfg, pv, prog, io=traces[(i + 1) % 24][2].io)    # another execution's I/O transcript
        assert not ok
    for i in (0, 7, 23):
        assert batch[i] == oracle.prove(cfg, *traces[i])


def test_error_paths_return_codes(gpu_ctx):
    """error behaviour of the boundary: bad shapes / values are ZKIR_ERR_ARG (-1), unconstrained rows ZKIR_ERR_AIR (-6);
    the context stays usable afterwards"""
    import ctypes as C
    from zkir_b200 import _ffi
    fres, cols, pv = fib_trace(30)
    cfg = zkir_b200.ProverConfig(num_queries=4, pow_bits=2)
    bad_pv = pv.copy(); bad_pv[1] = P                     # not canonical
    with pytest.raises(zkir_b200.RuntimeError) as ei:
        gpu_ctx.prove_columns(cols, bad_pv, cfg, program=fres)
    assert ei.value.code == -1
    lie = pv.copy(); lie[4] = 0; lie[2] = 5               # "did not halt" but claims an exit code
    with pytest.raises(zkir_b200.RuntimeError) as ei:
        gpu_ctx.prove_columns(cols, lie, cfg)
    assert ei.value.code == -1
    l = _ffi.lib()
    params = _ffi.Params(1, 4, 2, 72, 4)                 # the v1 shape
    proof, plen = C.c_void_p(), C.c_size_t()
    rc = l.zkir_b200_prove(gpu_ctx._h, C.byref(params), cols.ctypes.data, 10, pv.ctypes.data_as(_ffi.u32p), C.byref(proof), C.byref(plen))
    assert rc == -1 and b"width" in l.zkir_b200_last_error(gpu_ctx._h)
    fresh = zkir_b200.Context(0)                          # no program set: refused, not proven against an empty ROM
    try:
        with pytest.raises(zkir_b200.RuntimeError) as ei:
            fresh.prove_columns(cols, pv, cfg)
        assert ei.value.code == -1 and "program" in str(ei.value)
    finally:
        fresh.close()
    other = list(fres.program.code); other[2] = zkir_b200.encode("addi", 3, 0, imm=31)
    with pytest.raises(zkir_b200.RuntimeError) as ei:     # a trace of ANOTHER program: the ROM lookups cannot balance
        gpu_ctx.prove_columns(cols, pv, cfg, program=other)
    assert ei.value.code == -6 and "lookup" in str(ei.value)
    prog = zkir_b200.assemble("addi r1, r0, 3\nmul r2, r1, r1\nadd r10, r0, r0\necall\n")
    res = zkir_b200.VM(prog, [], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    # a MUL has no selector in the core table (the full profile proves it: tests/test_full_profile.py); the write-log path is core only
    for call in (lambda: gpu_ctx.prove_rows(res.rows(), cfg, profile="core"), lambda: gpu_ctx.prove_writelog(res.writelog(), cfg)):
        with pytest.raises(zkir_b200.RuntimeError) as ei:
            call()
        assert ei.value.code == -6 and "row 1" in str(ei.value)
    ok, why = zkir_b200.verify(gpu_ctx.prove_columns(cols, pv, cfg, program=fres), cfg, pv, fres)   # still healthy
    assert ok, why


def test_full_size_lde_linearity_and_interpolation(gpu_ctx):
    """BASELINE config 2 shape (2^20 rows): too large for the scalar oracle in CI time, so check size-independent properties
    of the same kernels the prover uses, bit-exact: LDE(a + b) = LDE(a) + LDE(b), and iNTT(NTT(a)) = a."""
    log_n, n_cols = 20, 6
    rng = np.random.default_rng(2020)
    a = rand_field(rng, (n_cols, 1 << log_n))
    b = rand_field(rng, (n_cols, 1 << log_n))
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    outs = []
    for m in (a, b, s):
        d_in = gpu_ctx.to_device(m)
        d_out = gpu_ctx.alloc(m.nbytes * 2)
        gpu_ctx.lde(d_in, d_out, n_cols, log_n, 1)
        outs.append(gpu_ctx.to_host(d_out, (n_cols, 2 << log_n)))
        gpu_ctx.free(d_in); gpu_ctx.free(d_out)
    la, lb, ls = outs
    assert np.array_equal(((la.astype(np.uint64) + lb) % P).astype(np.uint32), ls)
    # forward NTT then inverse NTT of the full-size matrix is the identity (natural order API, two-digit fast path)
    d = gpu_ctx.to_device(a)
    gpu_ctx.ntt(d, n_cols, log_n, inverse=False)
    gpu_ctx.ntt(d, n_cols, log_n, inverse=True)
    assert np.array_equal(gpu_ctx.to_host(d, a.shape), a)
    gpu_ctx.free(d)


def test_minimal_programs(gpu_ctx, oracle):
    """smallest traces: a lone ECALL (r10 = 0 -> EXIT 0 at cycle 0: one live row padded to 2^10) and a 2-cycle exit with a code"""
    cfg = zkir_b200.ProverConfig(num_queries=5, pow_bits=3)
    for src, code, cycles in (("ecall\n", 0, 1), ("addi r11, r0, 9\necall\n", 9, 2)):
        res = zkir_b200.VM(zkir_b200.assemble(src), [], zkir_b200.VMConfig(enable_execution_trace=True)).run()
        assert res.cycles == cycles and res.halt_reason == zkir_b200.HaltReason.Exit(code)
        cols, pv = res.pack()
        assert cols.shape[1] == 1024 and oracle.check_trace(cols, pv, res)[0] == -1
        want = oracle.prove(cfg, cols, pv, res)
        assert gpu_ctx.prove_columns(cols, pv, cfg, program=res) == want
        assert gpu_ctx.prove_rows(res.rows(), cfg)[0] == want
        assert gpu_ctx.prove_writelog(res.writelog(), cfg)[0] == want
        ok, why = zkir_b200.verify(want, cfg, pv, res)
        assert ok, why


def test_small_proofs_replay_a_captured_graph(gpu_ctx, oracle):
    """Proofs of up to 2^12 rows (every small program: 2^10 rows is the minimum) run as ONE captured CUDA graph from the third proof of a shape on (first: ordinary launches
    that fill the table caches, second: capture + replay).  Every replay must pick up the new trace / public values / PoW
    parameter and give the oracle's bytes."""
    cfg = zkir_b200.ProverConfig(num_queries=12, pow_bits=5)
    prog = zkir_b200.assemble(ADD_SRC)
    ctx = zkir_b200.Context(0)
    try:
        for i in range(6):
            res = zkir_b200.VM(prog, [7 * i + 1, i * i], zkir_b200.VMConfig(enable_execution_trace=True)).run()
            cols, pv = res.pack()
            if i == 4:
                cfg = zkir_b200.ProverConfig(num_queries=12, pow_bits=9)   # same workspace shape, different PoW: re-capture
            got = ctx.prove_columns(cols, pv, cfg, program=res)
            assert got == oracle.prove(cfg, cols, pv, res), f"proof {i} differs from the oracle"
        fres, cols, pv = fib_trace(205)                                    # another program on the same context and shape
        cfg = zkir_b200.ProverConfig(num_queries=20, pow_bits=6)
        want = oracle.prove(cfg, cols, pv, fres)
        for i in range(4):
            assert ctx.prove_columns(cols, pv, cfg, program=fres) == want
    finally:
        ctx.close()


def test_poseidon2_witness_of_a_run_recomputed_on_the_device(gpu_ctx):
    """BASELINE config 3 in miniature: the Poseidon2Witness records of a SYS_POSEIDON2 loop (zkir-spec/src/trace.rs:287-304) -- every
    permutation the interpreter executed -- recomputed in one batch by zkir_b200_poseidon2_permute"""
    from zkir_b200.workloads import pos2_program
    res = zkir_b200.VM(pos2_program(), [3000], zkir_b200.VMConfig(max_cycles=1 << 20, enable_execution_trace=True, enable_poseidon2_syscall=True)).run()
    ts, ins, outs = res.poseidon2_witness
    assert ins.shape == (3000, 16)
    d = gpu_ctx.to_device(ins)
    gpu_ctx.poseidon2_permute(d, ins.shape[0])
    got = gpu_ctx.to_host(d, ins.shape)
    gpu_ctx.free(d)
    assert np.array_equal(got, outs)
