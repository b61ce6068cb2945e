"""The FULL AIR profile (docs/PROVER_SPEC.md sections 3.7, 3.8): MUL MULH DIVU REMU DIV REM, AND OR XOR (+ immediates), the six shifts,
SLT SGE BLT BGE and the loads / stores (offline memory checking) on top of the core opcodes: all 50 opcodes of zkir-spec/src/opcode.rs.  CPU tests: the packer's witness against the oracle's row-by-row AIR check with expected
register values restated from zkir-runtime/src/execute.rs (:80-183 arithmetic, :200-279 logical, :282-358 shifts, :361-391 / :594-609
signed compares), tamper tests per family, oracle proof -> product verifier.  GPU tests: proof bytes == oracle, verifier accepts."""
import numpy as np
import pytest

import zkir_b200
from conftest import Oracle
from zkir_b200.runtime import FULL_WIDTH, program_profile
from zkir_b200.workloads import mix_program, mix_cycles, mix_reference
from zkir_b200 import air_layout_full

M40 = (1 << 40) - 1
LF = air_layout_full.INDEX


@pytest.fixture(scope="module")
def oracle_full():
    return Oracle(width=FULL_WIDTH)


def sg(v):
    return v - (1 << 40) if v >> 39 else v


def run(src, inputs=()):
    return zkir_b200.VM(zkir_b200.assemble(src), list(inputs), zkir_b200.VMConfig(enable_execution_trace=True)).run()


# every full-profile opcode on two operands from the input tape; the result lands in r3 and is written to the output tape
OPS = {
    "mul": lambda a, b: (a * b) & M40,
    "mulh": lambda a, b: ((a * b) >> 40) & M40,
    "divu": lambda a, b: a // b,
    "remu": lambda a, b: a % b,
    "div": lambda a, b: a // b,       # register values below 2^40 are non-negative as i64 (execute.rs:117-132)
    "rem": lambda a, b: a % b,
    "and": lambda a, b: a & b,
    "or": lambda a, b: a | b,
    "xor": lambda a, b: a ^ b,
    "sll": lambda a, b: (a << (b & 63)) & M40 if (b & 63) < 40 else 0,
    "srl": lambda a, b: a >> (b & 63) if (b & 63) < 40 else 0,
    "sra": lambda a, b: (sg(a) >> min(b & 63, 40)) & M40,
    "slt": lambda a, b: int(sg(a) < sg(b)),
    "sge": lambda a, b: int(sg(a) >= sg(b)),
}
PAIRS = [(0, 1), (1, 1), (12345, 7), (M40, M40), (M40, 1), (1 << 39, 3), (3, 1 << 39), ((1 << 39) + 5, (1 << 39) + 9), (0xABCDE12345, 0x1234567),
         (0x8000000001, 37), (0x7FFFFFFFFF, 39), (0xFFFFF00000, 40), (0xF0F0F0F0F0, 63), (0x123, 64 + 5), (0xFEDCBA9876, 1000), (5, 0xFFFFFFFFFF)]


def two_operand_program(ops):
    body = "addi r10, r0, 1\necall\nadd r1, r10, r0\naddi r10, r0, 1\necall\nadd r2, r10, r0\n"
    for op in ops:
        body += f"{op} r3, r1, r2\nadd r11, r3, r0\naddi r10, r0, 2\necall\n"
    return body + "addi r10, r0, 0\naddi r11, r0, 0\necall\n"


@pytest.mark.parametrize("a,b", PAIRS)
def test_register_forms_match_execute_rs_and_satisfy_the_air(oracle_full, a, b):
    ops = list(OPS)
    res = run(two_operand_program(ops), [a, b])
    assert res.outputs == [OPS[o](a, b) for o in ops], [(o, hex(x), hex(OPS[o](a, b))) for o, x in zip(ops, res.outputs) if x != OPS[o](a, b)]
    assert program_profile(res.program) == "full"
    cols, pv = res.pack()
    assert cols.shape[0] == FULL_WIDTH
    k, row = oracle_full.check_trace(cols, pv, res)
    assert k == -1, (k, row)


@pytest.mark.parametrize("a", [0, 5, 0xFFFFF, 0x8000000000, 0xFEDCBA9876, M40])
def test_immediate_forms_and_signed_branches(oracle_full, a):
    imms = [0, 1, 255, -1, -256, 65535, -65536]
    shs = [0, 1, 9, 10, 19, 20, 39, 40, 41, 63]
    src = "addi r10, r0, 1\necall\nadd r1, r10, r0\naddi r9, r0, 0\n"
    want = []
    for im in imms:
        for op, f in (("andi", lambda x, y: x & y), ("ori", lambda x, y: x | y), ("xori", lambda x, y: x ^ y)):
            src += f"{op} r3, r1, {im}\nadd r11, r3, r0\naddi r10, r0, 2\necall\n"
            want.append(f(a, im & M40))
    for sh in shs:
        src += f"slli r3, r1, {sh}\nadd r11, r3, r0\naddi r10, r0, 2\necall\nsrli r3, r1, {sh}\nadd r11, r3, r0\naddi r10, r0, 2\necall\nsrai r3, r1, {sh}\nadd r11, r3, r0\naddi r10, r0, 2\necall\n"
        want += [(a << sh) & M40 if sh < 40 else 0, a >> sh if sh < 40 else 0, (sg(a) >> min(sh, 40)) & M40]
    # signed branches against 0 and -1: r9 collects the taken ones
    src += "addi r2, r0, -1\n"
    for op, f, other in (("blt", lambda x, y: x < y, "r0"), ("bge", lambda x, y: x >= y, "r0"), ("blt", lambda x, y: x < y, "r2"), ("bge", lambda x, y: x >= y, "r2")):
        src += f"{op} r1, {other}, 8\naddi r9, r9, 1\n"
    src += "add r11, r9, r0\naddi r10, r0, 2\necall\naddi r10, r0, 0\naddi r11, r0, 0\necall\n"
    not_taken = sum(int(not f(sg(a), o)) for f, o in ((lambda x, y: x < y, 0), (lambda x, y: x >= y, 0), (lambda x, y: x < y, -1), (lambda x, y: x >= y, -1)))
    res = run(src, [a])
    assert res.outputs == want + [not_taken]
    cols, pv = res.pack()
    k, row = oracle_full.check_trace(cols, pv, res)
    assert k == -1, (k, row)


def test_tampered_full_profile_cells_are_rejected(oracle_full):
    a, b = 0xABCDE12345, 0x8000000025
    ops = list(OPS)
    res = run(two_operand_program(ops), [a, b])
    cols, pv = res.pack()
    assert oracle_full.check_trace(cols, pv, res)[0] == -1
    rows = res.rows()
    opnum = {"mul": 0x02, "mulh": 0x03, "divu": 0x04, "remu": 0x05, "div": 0x06, "rem": 0x07, "and": 0x10, "or": 0x11, "xor": 0x12, "sll": 0x18,
             "srl": 0x19, "sra": 0x1A, "slt": 0x22, "sge": 0x23}
    by_op = {}
    for i in range(res.cycles):
        by_op.setdefault(int(rows["instrs"][i]) & 0x7F, i)
    cases = [("mul", "v_lo"), ("mul", "p0"), ("mul", "p3"), ("mul", "k2_lo"), ("mul", "x1"), ("mul", "y2"), ("mulh", "v_hi"), ("mulh", "p6"), ("mulh", "k4_hi"),
             ("divu", "v_lo"), ("divu", "x0"), ("divu", "r0"), ("divu", "carry1"), ("remu", "v_lo"), ("remu", "r1"), ("div", "p4"), ("rem", "ch0"),
             ("and", "v_lo"), ("and", "zl0"), ("and", "xl1"), ("or", "v_hi"), ("or", "zh3"), ("xor", "v_lo"), ("xor", "yh2"),
             ("sll", "v_lo"), ("sll", "shamt"), ("sll", "sh_w"), ("sll", "y0"), ("srl", "v_lo"), ("srl", "y3"), ("srl", "sh_zero"), ("srl", "fill_hi"),
             ("sra", "v_hi"), ("sra", "sign_a"), ("sra", "fill_lo"), ("slt", "v_lo"), ("slt", "sign_a"), ("slt", "sign_b"), ("slt", "lt_signed"),
             ("slt", "sign_xor"), ("sge", "v_lo"), ("sge", "carry1"), ("mul", "m_rng"), ("and", "m_and"), ("sll", "m_pow")]
    for op, cell in cases:
        i = by_op[opnum[op]]
        bad = cols.copy()
        bad[LF[cell], i] ^= 1
        assert oracle_full.check_trace(bad, pv, res)[0] != -1, (op, cell)
    # a right-shift table row must not serve a left shift: move the key by 64 and compensate in w (tools/gen_air.py: key = shamt + 1024 right)
    i = by_op[0x18]
    bad = cols.copy()
    if int(bad[LF["sh_w"], i]) > 0:
        bad[LF["shamt"], i] += 64; bad[LF["sh_w"], i] -= 1
        assert oracle_full.check_trace(bad, pv, res)[0] != -1


def test_profiles_of_programs_and_what_no_profile_constrains():
    assert program_profile(zkir_b200.assemble("addi r1, r0, 3\nadd r2, r1, r1\nebreak")) == "core"
    for src in ("addi r1, r0, 3\nmul r2, r1, r1\nebreak", "addi r1, r0, 3\nand r2, r1, r1\nebreak", "addi r1, r0, 3\nslt r2, r1, r1\nebreak", "srai r1, r1, 3\nebreak"):
        res = run(src)
        assert program_profile(res.program) == "full"
        with pytest.raises(zkir_b200.RuntimeError) as e:   # the core table has no selector for these opcodes
            res.pack(profile="core")
        assert e.value.code == -6 and "not constrained" in str(e.value)
        assert res.pack()[0].shape[0] == FULL_WIDTH
    # loads and stores need the full profile's memory argument
    res = run("addi r1, r0, 0x2000\nsw r1, 0(r1)\nlw r2, 0(r1)\nebreak")
    assert program_profile(res.program) == "full"
    with pytest.raises(zkir_b200.RuntimeError) as e:
        res.pack(profile="core")
    assert e.value.code == -6
    # an immediate shift amount above 63 behaves like "40 or more" upstream (value.rs:658-691); the power table stops at 63
    res = run("addi r1, r0, 3\nslli r2, r1, 64\nebreak")
    assert int(res.rows()["final_regs"][2]) == 0
    with pytest.raises(zkir_b200.RuntimeError) as e:
        res.pack()
    assert "shift amount" in str(e.value)
    # a core program may also be proven with the full table
    res = run("addi r1, r0, 3\nadd r2, r1, r1\nebreak")
    assert res.pack(profile="full")[0].shape[0] == FULL_WIDTH


def test_mix_workload_and_oracle_proof_verifies(oracle_full):
    iters = 40
    res = zkir_b200.VM(mix_program(), [iters], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    assert res.cycles == mix_cycles(iters) and res.halt_reason == zkir_b200.HaltReason.Exit(mix_reference(iters))
    cols, pv = res.pack()
    assert oracle_full.check_trace(cols, pv, res) == (-1, 0)
    cfg = zkir_b200.ProverConfig(num_queries=12, pow_bits=4)
    pb = oracle_full.prove(cfg, cols, pv, res)
    assert zkir_b200.verify(pb, cfg, pv, res) == (True, "")
    lie = pv.copy(); lie[2] ^= 1
    assert not zkir_b200.verify(pb, cfg, lie, res)[0]
    other = zkir_b200.assemble(zkir_b200.workloads.MIX_SRC.replace("25173", "25171"))
    assert not zkir_b200.verify(pb, cfg, pv, other, io=res.io)[0], "a proof must not verify against another program"
    for off in (40, 4000, len(pb) // 2, len(pb) - 8):
        bad = bytearray(pb); bad[off] ^= 1
        assert not zkir_b200.verify(bytes(bad), cfg, pv, res)[0]
    # a core-profile proof of a core program still verifies next to it (the verifier dispatches on the header's width)
    r2 = run("addi r1, r0, 3\nadd r2, r1, r1\nebreak")
    c2, pv2 = r2.pack()
    assert zkir_b200.verify(Oracle().prove(cfg, c2, pv2, r2), cfg, pv2, r2) == (True, "")
    cf, pvf = r2.pack(profile="full")
    assert zkir_b200.verify(oracle_full.prove(cfg, cf, pvf, r2), cfg, pvf, r2) == (True, "")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("iters,log_b,nq", [(10, 1, 20), (200, 1, 30), (60, 2, 10)])
def test_gpu_full_profile_proof_equals_oracle(gpu_ctx, oracle_full, iters, log_b, nq):
    res = zkir_b200.VM(mix_program(), [iters], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    cols, pv = res.pack()
    cfg = zkir_b200.ProverConfig(log_blowup=log_b, num_queries=nq, pow_bits=6)
    gpu_ctx.set_io(res.io)
    pb = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    assert pb == oracle_full.prove(cfg, cols, pv, res), "GPU proof bytes differ from the CPU oracle (full profile)"
    assert zkir_b200.verify(pb, cfg, pv, res) == (True, "")


@pytest.mark.gpu
def test_gpu_full_profile_every_opcode(gpu_ctx, oracle_full):
    res = run(two_operand_program(list(OPS)), [0xABCDE12345, 0x8000000025])
    cols, pv = res.pack()
    cfg = zkir_b200.ProverConfig(num_queries=16, pow_bits=4)
    gpu_ctx.set_io(res.io)
    pb = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    assert pb == oracle_full.prove(cfg, cols, pv, res)
    assert zkir_b200.verify(pb, cfg, pv, res) == (True, "")
    # the device-side lookup balance check refuses a witness whose multiplier carry is wrong
    bad = cols.copy()
    i = next(i for i in range(res.cycles) if int(res.rows()["instrs"][i]) & 0x7F == 0x02)
    bad[LF["k1_lo"], i] = 5000
    with pytest.raises(zkir_b200.RuntimeError):
        gpu_ctx.prove_columns(bad, pv, cfg, program=res)


@pytest.mark.gpu
def test_gpu_full_profile_2p16_rows_and_prove_api(gpu_ctx, oracle_full):
    iters = 2900     # 63 811 cycles -> 2^16 rows
    prog = mix_program()
    cfg = zkir_b200.ProverConfig(num_queries=30, pow_bits=8)
    proof = zkir_b200.prove(prog, [iters], cfg)
    assert proof.log_n == 16 and proof.cycles == mix_cycles(iters) and int.from_bytes(proof.bytes_[12:16], "little") == FULL_WIDTH
    assert zkir_b200.verify(proof, cfg) == (True, "")
    res = zkir_b200.VM(prog, [iters], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    cols, pv = res.pack()
    assert proof.bytes_ == oracle_full.prove(cfg, cols, pv, res)


# ------------------------------------------------------------------------------------------------ loads and stores (section 3.8)
MEM_SRC = """
    addi r1, r0, 0x2000
    addi r2, r0, 1234
    sw r2, 0(r1)
    lw r3, 0(r1)
    addi r4, r0, -2
    sd r4, 8(r1)
    ld r5, 8(r1)
    sb r2, 3(r1)
    lbu r6, 3(r1)
    lb r7, 1(r1)
    sh r4, 6(r1)
    lhu r8, 6(r1)
    lh r9, 0(r1)
    lw r12, 4(r1)
    lw r13, 0x1000(r0)
    ld r14, -8(r1)
    addi r1, r1, 0x1000
    sw r2, 16(r1)
    lw r15, 16(r1)
    ebreak
"""


def test_loads_and_stores_satisfy_the_air(oracle_full):
    res = run(MEM_SRC)
    f = [int(x) for x in res.rows()["final_regs"]]
    # execute.rs:477-575: little endian, loads zero-extend (LW too: Appendix A), stores truncate to the width
    word0 = (1234 & 0x00FFFFFF) | ((1234 & 0xFF) << 24)          # sw 1234, then sb 0xD2 at byte 3
    assert f[3] == 1234 and f[5] == (-2) & M40 and f[6] == 0xD2 and f[7] == 0x04 and f[8] == 0xFFFE and f[9] == 0x04D2
    assert f[12] == 0xFFFE0000 and f[13] == res.program.code[0] and f[14] == 0 and f[15] == 1234
    assert word0 == 0xD20004D2
    cols, pv = res.pack()
    assert cols.shape[0] == FULL_WIDTH and oracle_full.check_trace(cols, pv, res) == (-1, 0)
    rows = res.rows()
    by_op = {}
    for i in range(res.cycles):
        by_op.setdefault(int(rows["instrs"][i]) & 0x7F, i)
    for op, cell in ((0x3A, "nb0"), (0x3A, "ob0"), (0x3A, "gb1"), (0x3A, "off0"), (0x3A, "mw"), (0x3A, "prev_ts"), (0x3A, "td0"), (0x3A, "ch1"), (0x3A, "nb5"),
                     (0x34, "v_lo"), (0x34, "ob1"), (0x34, "nb2"), (0x34, "gb0"), (0x34, "prev_ts"), (0x34, "nib_lo"), (0x35, "v_hi"), (0x35, "ob6"), (0x35, "gb4"),
                     (0x3B, "nb7"), (0x3B, "gb3"), (0x38, "nb3"), (0x38, "off3"), (0x31, "v_lo"), (0x30, "gb0"), (0x39, "nb6"), (0x39, "nb7"), (0x33, "gb1"),
                     (0x32, "v_lo"), (0x3A, "carry0"), (0x35, "carry1"), (0x3A, "m_b8"), (0x34, "m_b4"), (0x34, "m_b7")):
        bad = cols.copy()
        bad[LF[cell], by_op[op]] ^= 1
        assert oracle_full.check_trace(bad, pv, res)[0] != -1, (hex(op), cell)
    # boundary cells: final values / timestamps of image and RAM words, the RAM list itself
    n_img = int(zkir_b200._ffi.lib().zkir_image_words(len(res.program.code)))
    for cell, row in (("img_fin0", n_img - 1), ("img_fin_ts", 0x1000 // 8), ("ram_on", 0), ("ram_on", 3), ("ram_a0", 0), ("ram_a0", 1), ("ram_e0", 1), ("ram_fin0", 0),
                      ("ram_fin_ts", 1), ("ram_fin3", 2)):
        bad = cols.copy()
        bad[LF[cell], row] ^= 1
        assert oracle_full.check_trace(bad, pv, res)[0] != -1, (cell, row)
    # a word listed twice would fork memory: duplicate the first RAM entry onto the next free row
    nram = int(cols[LF["ram_on"]].sum())
    bad = cols.copy()
    for name in ["ram_on"] + [f"ram_a{k}" for k in range(3)] + [f"ram_fin{k}" for k in range(8)] + ["ram_fin_ts"]:
        bad[LF[name], nram] = bad[LF[name], 0]
    assert oracle_full.check_trace(bad, pv, res)[0] != -1


def test_reference_memory_and_arithmetic_programs_are_provable(oracle_full):
    """Programs of the reference's own tests: tests/stress_tests.rs:192-216 (100 stores over the program's own first words), :218-248
    (sparse addresses), :304-330 (add sub mul divu remu), tests/cross_module.rs:334-365 (store / load round trip, output [42])."""
    many = "addi r1, zero, 0x1000\naddi r2, zero, 1\n" + "".join(f"sw r2, {4 * i}(r1)\naddi r2, r2, 1\n" for i in range(100)) + "addi t2, zero, 0\naddi a0, zero, 0\necall\n"
    sparse = "addi r1, zero, 42\naddi r2, zero, 0x1000\nsw r1, 0(r2)\naddi r2, zero, 0x2000\nsw r1, 0(r2)\naddi r2, zero, 0x3000\nsw r1, 0(r2)\naddi t2, zero, 0\naddi a0, zero, 0\necall\n"
    arith = "addi r1, zero, 100\naddi r2, zero, 7\nadd r3, r1, r2\nsub r4, r1, r2\nmul r5, r1, r2\ndivu r6, r1, r2\nremu r7, r1, r2\naddi t2, zero, 0\naddi a0, zero, 0\necall\n"
    round_trip = "addi r1, zero, 42\naddi r2, zero, 0x1000\nsw r1, 0(r2)\nlw r3, 0(r2)\naddi a0, r3, 0\naddi t2, zero, 2\necall\naddi t2, zero, 0\naddi a0, zero, 0\necall\n"
    cfg = zkir_b200.ProverConfig(num_queries=8, pow_bits=2)
    for src, outputs in ((many, []), (sparse, []), (arith, []), (round_trip, [42])):
        res = run(src)
        assert res.halt_reason == zkir_b200.HaltReason.Exit(0) and res.outputs == outputs
        cols, pv = res.pack()
        assert oracle_full.check_trace(cols, pv, res) == (-1, 0)
        assert zkir_b200.verify(oracle_full.prove(cfg, cols, pv, res), cfg, pv, res) == (True, "")
    f = [int(x) for x in run(arith).rows()["final_regs"]]
    assert f[3:8] == [107, 93, 700, 14, 2]


def test_memory_rows_outside_the_model_are_rejected():
    # lb / lh of a negative value sign-extend to 64 bits upstream (execute.rs:477-500): outside the 40-bit register model
    res = run("addi r1, r0, 0x2000\naddi r2, r0, 200\nsb r2, 0(r1)\nlb r3, 0(r1)\nebreak")
    assert int(res.rows()["final_regs"][3]) == (200 - 256) & (2**64 - 1)
    with pytest.raises(zkir_b200.RuntimeError) as e:
        res.pack()
    assert e.value.code == -6
    # the unsigned form of the same load is fine
    res = run("addi r1, r0, 0x2000\naddi r2, r0, 200\nsb r2, 0(r1)\nlbu r3, 0(r1)\nebreak")
    assert res.pack()[0].shape[0] == FULL_WIDTH
    # SYS_POSEIDON2 writes guest memory outside the memory argument: a later load of its output is refused, not mis-proven
    src = "addi r11, r0, 0x2000\naddi r13, r0, 0x2000\naddi r10, r0, 4\necall\nlw r3, 0(r13)\nebreak"
    res = zkir_b200.VM(zkir_b200.assemble(src), [], zkir_b200.VMConfig(enable_execution_trace=True, enable_poseidon2_syscall=True)).run()
    with pytest.raises(zkir_b200.RuntimeError) as e:
        res.pack()
    assert "memory the AIR tracks" in str(e.value)


@pytest.mark.gpu
def test_gpu_memory_programs_equal_oracle(gpu_ctx, oracle_full):
    many = "addi r1, zero, 0x1000\naddi r2, zero, 1\n" + "".join(f"sw r2, {4 * i}(r1)\naddi r2, r2, 1\n" for i in range(100)) + "addi t2, zero, 0\naddi a0, zero, 0\necall\n"
    cfg = zkir_b200.ProverConfig(num_queries=16, pow_bits=4)
    for src in (MEM_SRC, many):
        res = run(src)
        cols, pv = res.pack()
        gpu_ctx.set_io(res.io)
        pb = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
        assert pb == oracle_full.prove(cfg, cols, pv, res)
        assert zkir_b200.verify(pb, cfg, pv, res) == (True, "")
        # Program -> Proof (interpreter write log + memory log -> device converter) and rows -> proof give the same bytes
        assert zkir_b200.prove(res.program, [], cfg).bytes_ == pb
        assert gpu_ctx.prove_rows(res.rows(), cfg)[0] == pb


@pytest.mark.gpu
def test_gpu_full_profile_from_write_log_and_memory_log(gpu_ctx, oracle_full):
    """zkir_b200_prove_writelog_mem: 28 B/row of logs -> the same proof bytes as the host packer + oracle; inconsistent logs never give a
    proof that verifies."""
    cfg = zkir_b200.ProverConfig(num_queries=16, pow_bits=4)
    for prog, inputs in ((mix_program(), [700]), (zkir_b200.assemble(MEM_SRC), [])):
        res = zkir_b200.VM(prog, inputs, zkir_b200.VMConfig(enable_execution_trace=True)).run()
        cols, pv = res.pack()
        want = oracle_full.prove(cfg, cols, pv, res)
        r = zkir_b200.VM(prog, inputs, zkir_b200.VMConfig(max_cycles=res.cycles + 8)).run_writelog()
        wl = r.writelog()
        assert "mem_old" in wl and wl["pcs"].shape[0] == res.cycles
        pb, pv2 = gpu_ctx.prove_writelog(wl, cfg)
        assert pb == want and np.array_equal(pv2, pv)
    # the register write log alone cannot describe a full-profile run
    core_wl = {k: v for k, v in wl.items() if not k.startswith("mem_")}
    with pytest.raises(zkir_b200.RuntimeError) as e:
        gpu_ctx.prove_writelog(core_wl, cfg)          # core width: the ROM lookup / opcode check refuses the rows
    import ctypes as C
    params = cfg.params(FULL_WIDTH)
    proof, plen = C.c_void_p(), C.c_size_t()
    pvb = np.zeros(5, dtype=np.uint32)
    rc = gpu_ctx._l.zkir_b200_prove_writelog(gpu_ctx._h, C.byref(params), wl["pcs"].ctypes.data, wl["instrs"].ctypes.data, wl["wlog"].ctypes.data, len(wl["pcs"]),
                                             int(wl["final_pc"]), int(wl["entry_point"]), 0, int(wl["halt_kind"]), 10, pvb.ctypes.data_as(zkir_b200._ffi.u32p),
                                             C.byref(proof), C.byref(plen))
    assert rc == -1 and b"memory log" in gpu_ctx._l.zkir_b200_last_error(gpu_ctx._h)
    # tampered logs: a wrong old word contradicts the loaded value (refused); a wrong previous timestamp or final word breaks the bus (rejected)
    rows = res.rows()
    ld = next(i for i in range(res.cycles) if int(rows["instrs"][i]) & 0x7F in (0x30, 0x31, 0x32, 0x33, 0x34, 0x35))
    bad = dict(wl); bad["mem_old"] = wl["mem_old"].copy(); bad["mem_old"][ld] ^= 0xFF
    with pytest.raises(zkir_b200.RuntimeError):
        gpu_ctx.prove_writelog(bad, cfg)
    for key, idx in (("mem_pts", ld), ("mem_word", 0), ("mem_ts", 0)):
        bad = dict(wl); bad[key] = wl[key].copy(); bad[key][idx] += 1
        try:
            pb_bad, pv_bad = gpu_ctx.prove_writelog(bad, cfg)
        except zkir_b200.RuntimeError:
            continue
        assert zkir_b200.verify(pb_bad, cfg, pv_bad, res)[0] is False, key
    bad = dict(wl); bad["mem_widx"] = wl["mem_widx"].copy(); bad["mem_widx"][-1] = 1 << 45     # a word index no address reaches
    try:
        pb_bad, pv_bad = gpu_ctx.prove_writelog(bad, cfg)
    except zkir_b200.RuntimeError:
        pass
    else:
        assert zkir_b200.verify(pb_bad, cfg, pv_bad, res)[0] is False
    pb, _ = gpu_ctx.prove_writelog(wl, cfg)           # the context is still good
    assert pb == want


@pytest.mark.gpu
def test_gpu_full_profile_program_to_proof_malformed_runs(gpu_ctx):
    """Rows the full profile cannot constrain must fail with an error through Program -> Proof too (the device converter reports them)."""
    cfg = zkir_b200.ProverConfig(num_queries=8, pow_bits=2)
    for src in ("addi r1, zero, 5\ndivu r2, r1, zero\nebreak\n",                    # division by zero
                "addi r1, zero, 0x1001\nlw r2, 0(r1)\nebreak\n"):                   # misaligned load: the interpreter refuses (memory.rs alignment)
        with pytest.raises(Exception):
            zkir_b200.prove(zkir_b200.assemble(src), [], cfg)
    # a store above 2^30 runs fine in the interpreter; the DEVICE converter reports the row (PACK_ERR_MEMADDR), as the host packer does
    with pytest.raises(zkir_b200.RuntimeError) as e:
        zkir_b200.prove(zkir_b200.assemble("addi r1, r0, 1\nslli r1, r1, 30\nsd r1, 0(r1)\nebreak"), [], cfg)
    assert e.value.code == -6 and "row 2" in str(e.value) and "memory address" in str(e.value)
    # far outside: the touched word has no place in the boundary's chunk tables either (it is left out there; the row is still reported)
    with pytest.raises(zkir_b200.RuntimeError) as e:
        zkir_b200.prove(zkir_b200.assemble("addi r1, r0, 1\nslli r1, r1, 39\nsd r1, 0(r1)\nebreak"), [], cfg)
    assert e.value.code == -6 and "row 2" in str(e.value) and "memory address" in str(e.value)
    # and the context stays usable
    assert zkir_b200.verify(zkir_b200.prove(mix_program(), [10], cfg), cfg) == (True, "")


@pytest.mark.gpu
@pytest.mark.parametrize("shards,min_seg", [(2, 2), (8, 2), (4, 0)])
def test_gpu_full_profile_emulated_shards_give_single_gpu_bytes(gpu_ctx, shards, min_seg):
    """the sharded code path (segment hashing, column / row / plane partitions: tests/test_gpu_sharded.py) takes the widths at run time"""
    res = zkir_b200.VM(mix_program(), [180], zkir_b200.VMConfig(enable_execution_trace=True)).run()   # 3971 cycles -> 2^12 rows
    cols, pv = res.pack()
    cfg = zkir_b200.ProverConfig(num_queries=24, pow_bits=6)
    gpu_ctx.set_io(res.io)
    want = gpu_ctx.prove_columns(cols, pv, cfg, program=res)
    ctx = zkir_b200.Context(0)
    try:
        ctx.emulate_shards(shards, min_seg)
        ctx.set_io(res.io)
        got = ctx.prove_columns(cols, pv, cfg, program=res)
    finally:
        ctx.close()
    assert got == want


def test_full_profile_golden_proof_digest(oracle_full):
    """Pins the full-profile generator, packer and oracle against drift (tests/golden/make_golden.py; self-generated: SURVEY.md 8c)."""
    import hashlib, json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
    res = zkir_b200.VM(mix_program(), [40], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    cols, pv = res.pack()
    pb = oracle_full.prove(zkir_b200.ProverConfig(num_queries=20, pow_bits=8), cols, pv, res)
    assert hashlib.sha256(cols.tobytes()).hexdigest() == gold["mix40_full_trace_sha256"]
    assert hashlib.sha256(pb).hexdigest() == gold["mix40_full_proof_sha256"] and len(pb) == gold["mix40_full_proof_len"]


@pytest.mark.gpu
def test_gpu_full_profile_golden_proof_digest(gpu_ctx):
    import hashlib, json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
    res = zkir_b200.VM(mix_program(), [40], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    cols, pv = res.pack()
    gpu_ctx.set_io(res.io)
    pb = gpu_ctx.prove_columns(cols, pv, zkir_b200.ProverConfig(num_queries=20, pow_bits=8), program=res)
    assert hashlib.sha256(pb).hexdigest() == gold["mix40_full_proof_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mem", "ops", "mix", "many_stores"])
def test_gpu_full_profile_device_converter_equals_host_packer(gpu_ctx, which):
    """zkir_b200_expand_rows_full (host memory replay + device row expansion, csrc/trace_expand.cu) == zkir_pack_trace_full, column for column"""
    if which == "mem":
        res = run(MEM_SRC)
    elif which == "ops":
        res = run(two_operand_program(list(OPS)), [0xABCDE12345, 0x8000000025])
    elif which == "mix":
        res = zkir_b200.VM(mix_program(), [700], zkir_b200.VMConfig(enable_execution_trace=True)).run()    # 2^14 rows
    else:
        res = run("addi r1, zero, 0x1000\naddi r2, zero, 1\n" + "".join(f"sw r2, {4 * i}(r1)\naddi r2, r2, 1\n" for i in range(100)) + "addi t2, zero, 0\naddi a0, zero, 0\necall\n")
    want, pv = res.pack()
    log_n = int(want.shape[1]).bit_length() - 1
    d = gpu_ctx.alloc(want.nbytes)
    try:
        gpu_ctx.expand_rows(res.rows(), log_n, d, profile="full")
        got = gpu_ctx.to_host(d, want.shape)
    finally:
        gpu_ctx.free(d)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, [air_layout_full.COLUMNS[c] for c in bad[:10]]


@pytest.mark.gpu
def test_gpu_full_profile_prove_rows_errors(gpu_ctx):
    cfg = zkir_b200.ProverConfig(num_queries=8, pow_bits=2)
    # a negative lb leaves the 40-bit register model: refused by the replay, like the host packer
    res = run("addi r1, r0, 0x2000\naddi r2, r0, 200\nsb r2, 0(r1)\nlb r3, 0(r1)\nebreak")
    with pytest.raises(zkir_b200.RuntimeError) as e:
        gpu_ctx.prove_rows(res.rows(), cfg, profile="full")
    assert e.value.code == -6 and "lb / lh" in str(e.value)
    # an immediate shift amount above 63: refused by the device converter
    res = run("addi r1, r0, 3\nslli r2, r1, 64\nebreak")
    with pytest.raises(zkir_b200.RuntimeError) as e:
        gpu_ctx.prove_rows(res.rows(), cfg, profile="full")
    assert e.value.code == -6 and "shift amount" in str(e.value)


def _run_memlog(prog, inputs, cap, poseidon2=False):
    import ctypes as C
    l = zkir_b200._ffi.lib()
    code = np.asarray(prog.code, dtype=np.uint32)
    arr = (np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint32))
    data = (C.c_uint8 * max(1, len(prog.data)))(*prog.data)
    inp = (C.c_uint64 * max(1, len(inputs)))(*inputs)
    h = C.c_void_p()
    l.zkir_vm_enable_poseidon2(int(poseidon2))
    rc = l.zkir_vm_run_writelog_mem_cb(code.ctypes.data_as(C.POINTER(C.c_uint32)), len(code), data, len(prog.data), prog.entry_point, inp, len(inputs), cap,
                                       *[a.ctypes.data for a in arr], cap, None, None, 0, C.byref(h))
    l.zkir_vm_enable_poseidon2(0)
    return rc, h, arr


def test_interpreter_memory_log_refuses_what_the_replay_refuses():
    """Same refusals as test_memory_rows_outside_the_model_are_rejected, through the memory-log interpreter (ZKIR_ERR_AIR = -6)."""
    l = zkir_b200._ffi.lib()
    asm = zkir_b200.assemble
    # a sign-extended load leaves the 40-bit register model
    rc, h, _ = _run_memlog(asm("addi r1, r0, 0x2000\naddi r2, r0, 200\nsb r2, 0(r1)\nlb r3, 0(r1)\nebreak"), [], 64)
    assert rc == -6 and b"above 40 bits" in l.zkir_vm_last_error()
    # SYS_POSEIDON2 writes memory the argument does not see: a later load (or store) of those words is refused ...
    pos = "addi r11, r0, 0x2000\naddi r13, r0, 0x2000\naddi r10, r0, 4\necall\n"
    for tail in ("lw r3, 0(r13)\nebreak", "sb r3, 60(r13)\nebreak", "lbu r3, 57(r13)\nebreak"):
        rc, h, _ = _run_memlog(asm(pos + tail), [], 64, poseidon2=True)
        assert rc == -6 and b"memory the AIR tracks" in l.zkir_vm_last_error(), tail
    # ... a neighbouring word is fine, and a word the program stored BEFORE the syscall overwrote it keeps the value the argument saw
    rc, h, arr = _run_memlog(asm("addi r13, r0, 0x2000\naddi r2, r0, 77\nsw r2, 4(r13)\naddi r11, r0, 0x2000\naddi r10, r0, 4\necall\nlw r3, 64(r13)\nebreak"), [], 64, poseidon2=True)
    assert rc == 0
    nw = l.zkir_vm_memlog_count(h)
    widx = np.ctypeslib.as_array(l.zkir_vm_memlog_widx(h), shape=(nw,)); word = np.ctypeslib.as_array(l.zkir_vm_memlog_word(h), shape=(nw,))
    assert list(widx) == [0x2000 // 8, 0x2040 // 8] and list(word) == [77 << 32, 0]
    l.zkir_vm_free(h)
    # a data segment is not part of the public image: its first load is refused
    prog = asm("addi r1, r0, 0x1000\nlw r2, 0(r1)\nlw r3, 12(r1)\nebreak")
    rc, h, _ = _run_memlog(prog, [], 64)
    assert rc == 0
    l.zkir_vm_free(h)
    prog.data = bytes([1, 2, 3, 4])
    rc, h, _ = _run_memlog(prog, [], 64)
    assert rc == 0   # never loaded
    l.zkir_vm_free(h)
    prog2 = asm("addi r1, r0, 0x1000\nlw r2, 16(r1)\nebreak\nebreak")
    prog2.data = bytes([1, 2, 3, 4])
    rc, h, _ = _run_memlog(prog2, [], 64)
    assert rc == -6 and b"data segment" in l.zkir_vm_last_error()


def test_interpreter_memory_log_equals_the_host_replay():
    """Program -> Proof for the full profile feeds the device converter from the interpreter's own memory log (zkir_vm_run_writelog_mem_cb:
    the word before each load / store, that word's previous timestamp, and the touched words at the end).  It must be the same data the
    host replay of recorded rows (zkir_mem_replay_full / zkir_mem_boundary_full, the prove_rows path) derives."""
    import ctypes as C
    from zkir_b200 import _ffi
    l = _ffi.lib()
    raw = C.CDLL(_ffi.LIB_PATH)   # the two-step host replay is internal to the library (not part of include/zkir_b200.h)
    import test_fuzz_full_profile as fz
    cases = [(mix_program(), [60]), (run(MEM_SRC).program, [])]
    for seed in (3, 11, 29):
        lines, body, inputs = fz.random_program(np.random.default_rng(1000 + seed), 200)
        cases.append((zkir_b200.assemble(fz.run_model_and_emit(lines, body, inputs)[0]), inputs))
    for prog, inputs in cases:
        res = zkir_b200.VM(prog, inputs, zkir_b200.VMConfig(enable_execution_trace=True)).run()
        rows = res.rows()
        T = len(rows["instrs"])
        code = np.asarray(prog.code, dtype=np.uint32)
        # (1) host replay of the recorded rows
        old_r = np.zeros(T, dtype=np.uint64); pts_r = np.zeros(T, dtype=np.uint32)
        n_img = C.c_uint64(); n_ram = C.c_uint64()
        regs = np.ascontiguousarray(rows["regs"]); fin = np.ascontiguousarray(rows["final_regs"]); ins = np.ascontiguousarray(rows["instrs"])
        raw.zkir_mem_replay_full.restype = C.c_int
        rc = raw.zkir_mem_replay_full(C.c_void_p(ins.ctypes.data), C.c_void_p(regs.ctypes.data), C.c_uint64(T), C.c_void_p(fin.ctypes.data),
                                      C.c_void_p(code.ctypes.data), C.c_size_t(len(code)), C.c_void_p(old_r.ctypes.data), C.c_void_p(pts_r.ctypes.data),
                                      C.byref(n_img), C.byref(n_ram))
        assert rc == 0
        log_n = 10
        while (1 << log_n) <= T:
            log_n += 1
        bs = max(n_img.value, n_ram.value, 1)
        b_r = np.zeros((25, bs), dtype=np.uint32); rng_r = np.zeros(1024, dtype=np.uint32); b7_r = np.zeros(128, dtype=np.uint32)
        rc = raw.zkir_mem_boundary_full(C.c_uint32(log_n), C.c_void_p(b_r.ctypes.data), C.c_uint64(bs), C.c_void_p(rng_r.ctypes.data), C.c_void_p(b7_r.ctypes.data))
        assert rc == 0
        # (2) the interpreter's memory log
        rc, h, (pcs, ins2, wlog, old_l, pts_l) = _run_memlog(prog, inputs, T + 8)
        assert rc == 0 and l.zkir_vm_logged_rows(h) == T
        assert np.array_equal(ins2[:T], ins)
        is_mem = (pts_r != 0) | (old_r != 0) | (pts_l[:T] != 0) | (old_l[:T] != 0)
        assert is_mem.any()
        assert np.array_equal(old_l[:T], old_r) and np.array_equal(pts_l[:T], pts_r)
        nw = l.zkir_vm_memlog_count(h)
        widx = np.ctypeslib.as_array(l.zkir_vm_memlog_widx(h), shape=(nw,)).copy()
        word = np.ctypeslib.as_array(l.zkir_vm_memlog_word(h), shape=(nw,)).copy()
        ts = np.ctypeslib.as_array(l.zkir_vm_memlog_ts(h), shape=(nw,)).copy()
        assert np.all(np.diff(widx.astype(np.int64)) > 0)   # ascending, no duplicates
        assert int((widx >= l.zkir_image_words(len(code))).sum()) == n_ram.value
        b_l = np.zeros((25, bs), dtype=np.uint32); rng_l = np.zeros(1024, dtype=np.uint32); b7_l = np.zeros(128, dtype=np.uint32)
        raw.zkir_mem_boundary_from_words_full.restype = C.c_int
        rc = raw.zkir_mem_boundary_from_words_full(C.c_void_p(code.ctypes.data), C.c_size_t(len(code)), C.c_void_p(widx.ctypes.data), C.c_void_p(word.ctypes.data),
                                                   C.c_void_p(ts.ctypes.data), C.c_size_t(nw), C.c_uint32(log_n), C.c_void_p(b_l.ctypes.data), C.c_uint64(bs),
                                                   C.c_void_p(rng_l.ctypes.data), C.c_void_p(b7_l.ctypes.data))
        assert rc == 0
        assert np.array_equal(b_l, b_r) and np.array_equal(rng_l, rng_r) and np.array_equal(b7_l, b7_r)
        l.zkir_vm_free(h)


def test_full_profile_verifier_survives_malformed_proofs(oracle_full):
    """Untrusted bytes against the full profile's verifier: extreme header fields (incl. the width word that selects the profile), a core
    proof offered for a full-profile program and the reverse are all rejected without a crash."""
    from zkir_b200.workloads import fib_trace
    res = zkir_b200.VM(mix_program(), [20], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    cols, pv = res.pack()
    cfg = zkir_b200.ProverConfig(num_queries=6, pow_bits=2)
    pb = oracle_full.prove(cfg, cols, pv, res)
    assert zkir_b200.verify(pb, cfg, pv, res) == (True, "")
    w = np.frombuffer(pb, dtype=np.uint32)
    for pos in range(24):
        for v in (0, 1, 7, 31, 32, 64, 88, 248, 255, 1 << 16, 1 << 20, 0x7FFFFFFF, 0xFFFFFFFF):
            if int(w[pos]) == v:
                continue
            bad = w.copy()
            bad[pos] = v
            assert not zkir_b200.verify(bad.tobytes(), cfg, pv, res)[0], (pos, v)
    r2, c2, p2 = fib_trace(30)
    core = Oracle().prove(cfg, c2, p2, r2.program)
    assert zkir_b200.verify(core, cfg, p2, r2.program) == (True, "")
    assert not zkir_b200.verify(core, cfg, p2, res)[0]            # a core proof is no statement about a program that needs the full profile
    assert not zkir_b200.verify(pb, cfg, pv, r2.program)[0]
    relabel = np.frombuffer(core, dtype=np.uint32).copy()
    relabel[3] = FULL_WIDTH
    assert not zkir_b200.verify(relabel.tobytes(), cfg, p2, r2.program)[0]
