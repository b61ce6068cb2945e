"""The FULL AIR profile (docs/PROVER_SPEC.md section 3.7): MUL MULH DIVU REMU DIV REM, AND OR XOR (+ immediates), the six shifts,
SLT SGE BLT BGE on top of the core opcodes.  CPU tests: the packer's witness against the oracle's row-by-row AIR check with expected
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
    # loads and stores: no memory argument in either profile
    res = run("addi r1, r0, 0x2000\nsw r1, 0(r1)\nlw r2, 0(r1)\nebreak")
    assert program_profile(res.program) is None
    for prof in ("core", "full"):
        with pytest.raises(zkir_b200.RuntimeError) as e:
            res.pack(profile=prof)
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
    iters = 4000     # 64 011 cycles -> 2^16 rows
    prog = mix_program()
    proof = zkir_b200.prove(prog, [iters], zkir_b200.ProverConfig(num_queries=30, pow_bits=8))
    assert proof.log_n == 16 and proof.cycles == mix_cycles(iters) and int.from_bytes(proof.bytes_[12:16], "little") == FULL_WIDTH
    assert zkir_b200.verify(proof) == (True, "")
    res = zkir_b200.VM(prog, [iters], zkir_b200.VMConfig(enable_execution_trace=True)).run()
    cols, pv = res.pack()
    assert proof.bytes_ == oracle_full.prove(zkir_b200.ProverConfig(num_queries=30, pow_bits=8), cols, pv, res)
