"""The upstream (VM / trace) half IS pinned by the reference's own tests.  Each test here restates an assertion of
seceq/zkir's test-suite (file:line cited) against the C++ interpreter behind the C ABI (zkir_b200/csrc/host/vm.cc)."""
import numpy as np
import pytest

import zkir_b200 as z
from zkir_b200 import encode, decode, assemble, Program, VM, VMConfig, HaltReason
from conftest import FIB_SRC


def prog(words):
    return Program(words)


def test_encode_bit_layout():
    # zkir-assembler/src/encoder.rs:158-230
    w = encode("add", 4, 5, 6)
    assert (w & 0x7F, (w >> 7) & 0xF, (w >> 11) & 0xF, (w >> 15) & 0xF) == (0x00, 4, 5, 6)
    w = encode("addi", 4, 5, imm=100)
    assert (w & 0x7F, (w >> 7) & 0xF, (w >> 11) & 0xF, (w >> 15) & 0x1FFFF) == (0x08, 4, 5, 100)
    assert encode("and", 2, 3, 4) & 0x7F == 0x10
    assert encode("lw", 4, 2, imm=16) & 0x7F == 0x34
    w = encode("beq", 4, 5, imm=8)
    assert (w & 0x7F, (w >> 7) & 0xF, (w >> 11) & 0xF, (w >> 15) & 0x1FFFF) == (0x40, 4, 5, 8)
    w = encode("jal", 1, imm=-8)
    assert w & 0x7F == 0x48 and (w >> 11) == ((-8) & 0x1FFFFF)
    assert encode("ecall") == 0x50 and encode("ebreak") == 0x51


@pytest.mark.parametrize("imm", [0, 1, -1, 127, -128, 255, -256, 32767, -32768, 65535, -65536])
def test_encode_decode_roundtrip_edge_immediates(imm):
    # tests/cross_module.rs:228-256
    for m in ("addi", "andi", "lw", "jalr"):
        assert decode(encode(m, 3, 7, imm=imm)) == (m, 3, 7, 0, imm)
    for m in ("sw", "beq", "bgeu"):
        assert decode(encode(m, 3, 7, imm=imm)) == (m, 3, 7, 0, imm)
    assert decode(encode("jal", 5, imm=imm)) == ("jal", 5, 0, 0, imm)


def test_roundtrip_all_opcodes_and_unknown():
    for m, op in z.OPCODES.items():
        assert decode(encode(m, 1, 2, 3, 4))[0] == m
    with pytest.raises(ValueError):
        decode(0x7F)   # zkir-disassembler/src/decoder.rs:24-25 UnknownOpcode
    # immediates outside 17 bits are silently truncated (encoder.rs:117)
    assert decode(encode("addi", 1, 0, imm=(1 << 17) + 5))[4] == 5
    assert decode(encode("slli", 1, 2, imm=39)) == ("slli", 1, 2, 0, 39)


def test_program_bytes_roundtrip_and_validation():
    # zkir-spec/src/program.rs:170-213,300-346
    p = Program([encode("addi", 1, 0, imm=5), encode("ebreak")], b"\x01\x02\x03")
    b = p.to_bytes()
    assert b[:4] == b"ZKIR" and len(b) == 32 + 8 + 3
    q = Program.from_bytes(b)
    assert q.code == p.code and q.data == p.data and q.entry_point == 0x1000
    with pytest.raises(ValueError):
        Program.from_bytes(b"XKIR" + b[4:])
    with pytest.raises(ValueError):
        Program.from_bytes(b[:-1])


def test_cycle_counts_and_halt_reasons():
    # zkir-runtime/src/vm.rs:433-486: 3 addi + ebreak = 4 cycles
    r = VM(prog([encode("addi", 1, 0, imm=10), encode("addi", 2, 0, imm=20), encode("add", 3, 1, 2), encode("ebreak")])).run()
    assert r.cycles == 4 and r.halt_reason == HaltReason.Ebreak
    # exit syscall: addi r10,0 ; addi r11,42 ; ecall -> Exit(42) in 3 cycles
    r = VM(prog([encode("addi", 10, 0, imm=0), encode("addi", 11, 0, imm=42), encode("ecall")])).run()
    assert r.cycles == 3 and r.halt_reason == HaltReason.Exit(42)
    # cycle limit (vm.rs:211-214, 536-552): infinite loop stops at max_cycles with Ok
    r = VM(prog([encode("jal", 0, imm=0)]), [], VMConfig(max_cycles=100)).run()
    assert r.cycles == 100 and r.halt_reason == HaltReason.CycleLimit


def test_1000_instructions():
    # tests/stress_tests.rs:24-56
    words = [encode("add", 1, 1, 0)] * 1000 + [encode("addi", 10, 0, imm=0), encode("addi", 11, 0, imm=0), encode("ecall")]
    r = VM(prog(words)).run()
    assert r.halt_reason == HaltReason.Exit(0) and r.cycles == 1003


def test_io_echo_and_sum():
    # tests/cross_module.rs:31-57 (outputs == [123]) using the assembler's aliases t2=r10, a0=r11
    src = """
        addi t2, zero, 1    # read
        ecall
        addi a0, t2, 0
        addi t2, zero, 2    # write
        ecall
        addi t2, zero, 0
        addi a0, zero, 0
        ecall
    """
    r = VM(assemble(src), [123]).run()
    assert r.outputs == [123] and r.halt_reason == HaltReason.Exit(0)
    # tests/cross_module.rs:59-86: 10 + 20 + 30 = 60
    src = """
        addi r1, zero, 10
        addi r2, zero, 20
        addi r3, zero, 30
        add r4, r1, r2
        add r4, r4, r3
        addi a0, r4, 0
        addi t2, zero, 2
        ecall
        addi t2, zero, 0
        addi a0, zero, 0
        ecall
    """
    assert VM(assemble(src)).run().outputs == [60]


def test_many_io_operations_and_edge_cases():
    # tests/stress_tests.rs:396-430
    src = """
        addi r3, zero, 5
    loop:
        addi t2, zero, 1
        ecall
        addi a0, t2, 0
        addi t2, zero, 2
        ecall
        addi r3, r3, -1
        bne r3, zero, -24
        addi t2, zero, 0
        addi a0, zero, 0
        ecall
    """
    assert VM(assemble(src), [1, 2, 3, 4, 5]).run().outputs == [1, 2, 3, 4, 5]
    # :436-460 division by one
    src = "addi r1, zero, 12345\naddi r2, zero, 1\ndivu r3, r1, r2\naddi a0, r3, 0\naddi t2, zero, 2\necall\naddi t2, zero, 0\naddi a0, zero, 0\necall"
    assert VM(assemble(src)).run().outputs == [12345]
    # :462-494 rd == rs1 == rs2
    src = "addi r1, zero, 10\nadd r1, r1, r1\nadd r1, r1, r1\nadd r1, r1, r1\naddi a0, r1, 0\naddi t2, zero, 2\necall\naddi t2, zero, 0\naddi a0, zero, 0\necall"
    assert VM(assemble(src)).run().outputs == [80]
    # :496-519 writes to the zero register are ignored
    src = "addi zero, zero, 100\naddi a0, zero, 0\naddi t2, zero, 2\necall\naddi t2, zero, 0\naddi a0, zero, 0\necall"
    assert VM(assemble(src)).run().outputs == [0]
    # read past the end of the tape returns 0 (syscall.rs:54-62)
    src = "addi t2, zero, 1\necall\naddi a0, t2, 0\naddi t2, zero, 2\necall\naddi t2, zero, 0\necall"
    assert VM(assemble(src), []).run().outputs == [0]


def test_fibonacci_runs_like_the_reference():
    # tests/end_to_end.rs:310-332 (n = 10): 4 + 5*(n-2) + 2 cycles, Exit(0)
    r = VM(assemble(FIB_SRC.format(n=10))).run()
    assert r.cycles == 46 and r.halt_reason == HaltReason.Exit(0)


def test_trace_shape_and_memory_ops():
    # zkir-runtime/src/vm.rs:906-964: 4 rows, cycles 0..3, 16 registers, pre-state
    cfg = VMConfig(enable_execution_trace=True)
    r = VM(prog([encode("addi", 1, 0, imm=10), encode("addi", 2, 0, imm=20), encode("add", 3, 1, 2), encode("ebreak")]), [], cfg).run()
    t = r.execution_trace
    assert len(t) == 4 and [row.cycle for row in t] == [0, 1, 2, 3] and all(len(row.registers) == 16 for row in t)
    assert t[0].pc == 0x1000 and t[1].pc == 0x1004 and t[0].registers[1] == 0 and t[1].registers[1] == 10 and t[3].registers[3] == 30
    assert t[2].instruction == encode("add", 3, 1, 2)
    # vm.rs:995-1075: sw to 0x1000 then lw; the store overwrites code (protection is off, vm.rs:175)
    words = [encode("addi", 1, 0, imm=0x42), encode("addi", 3, 0, imm=0x1000), encode("sw", 3, 1, imm=0), encode("lw", 4, 3, imm=0), encode("ebreak")]
    r = VM(prog(words), [], cfg).run()
    t = r.execution_trace
    assert r.halt_reason == HaltReason.Ebreak and len(t) == 5
    assert len(t[2].memory_ops) == 1 and t[2].memory_ops[0].is_write and t[2].memory_ops[0].width == 4
    assert len(t[3].memory_ops) == 1 and not t[3].memory_ops[0].is_write and t[3].memory_ops[0].value == 0x42
    assert len(t[0].memory_ops) == 0
    # timestamps equal the cycle of the row (vm.rs:1077-1200); sorted memory trace order trace.rs:210-223
    for row in t:
        assert all(op.timestamp == row.cycle for op in row.memory_ops)
    assert [op.timestamp for op in r.get_memory_trace()] == [2, 3] and r.memory_op_count() == 2
    # tracing off: no rows (vm.rs:25-31)
    assert VM(prog(words)).run().trace_len == 0


def test_semantics_40_bit_quirks():
    # SURVEY.md Appendix A restated from execute.rs / value.rs
    def out_of(src, inputs=()):
        return VM(assemble(src + "\naddi a0, r5, 0\naddi t2, zero, 2\necall\naddi t2, zero, 0\naddi a0, zero, 0\necall"), list(inputs)).run().outputs[0]
    M40 = (1 << 40) - 1
    assert out_of("addi r1, zero, -1\naddi r5, r1, 0") == M40                    # imm as u64 masked to 40 bits (execute.rs:187)
    assert out_of("addi r1, zero, -1\nadd r5, r1, r1") == (M40 + M40) & M40      # wrap mod 2^40 (value.rs:620-623)
    assert out_of("addi r1, zero, 1\nsub r5, zero, r1") == M40
    assert out_of("addi r1, zero, -1\nmulh r5, r1, r1") == ((M40 * M40) >> 40) & M40   # execute.rs:101-106
    assert out_of("addi r1, zero, 1\naddi r2, zero, 39\nsll r5, r1, r2") == 1 << 39
    assert out_of("addi r1, zero, 1\naddi r2, zero, 40\nsll r5, r1, r2") == 0   # shift >= 40 gives 0 (value.rs:658-663)
    assert out_of("addi r1, zero, -1\nsrai r5, r1, 45") == M40                   # sign fill
    assert out_of("addi r1, zero, -1\naddi r2, zero, 1\nslt r5, r1, r2") == 1    # signed at bit 39
    assert out_of("addi r1, zero, -1\naddi r2, zero, 1\nsltu r5, r1, r2") == 0
    assert out_of("addi r1, zero, 7\naddi r2, zero, 0\ncmov r5, r1, r2") == 0    # CMOV == CMOVNZ (execute.rs:434-474)
    assert out_of("addi r1, zero, 7\naddi r2, zero, 0\ncmovz r5, r1, r2") == 7
    # LB sign-extends to 64 bits, LW zero-extends (execute.rs:477-546)
    # (read r5 back through CMOV, which copies the raw 64-bit register; ADDI would mask it to 40 bits)
    raw = VM(assemble("addi r1, zero, 255\naddi r3, zero, 8192\nsb r1, 0(r3)\nlb r5, 0(r3)\naddi r2, zero, 1\ncmov a0, r5, r2\n"
                      "addi t2, zero, 2\necall\naddi t2, zero, 0\naddi a0, zero, 0\necall")).run().outputs[0]
    assert raw == (1 << 64) - 1
    assert out_of("addi r1, zero, -1\naddi r3, zero, 8192\nsw r1, 0(r3)\nlw r5, 0(r3)") == 0xFFFFFFFF


def test_runtime_errors():
    with pytest.raises(z.RuntimeError, match="Division by zero"):
        VM(assemble("addi r1, zero, 5\ndivu r2, r1, zero\nebreak")).run()       # execute.rs:117-183
    with pytest.raises(z.RuntimeError, match="Misaligned"):
        VM(assemble("addi r3, zero, 8193\nlw r1, 0(r3)\nebreak")).run()         # memory.rs:383
    with pytest.raises(z.RuntimeError, match="Invalid syscall"):
        VM(assemble("addi t2, zero, 99\necall")).run()                          # syscall.rs:173-175
    with pytest.raises(z.RuntimeError, match="Poseidon2 not yet implemented"):
        VM(assemble("addi t2, zero, 4\necall")).run()                           # crypto.rs:306-315, syscall_integration.rs:400-422
    with pytest.raises(z.RuntimeError, match="debug format"):
        VM(Program([encode("ebreak")], entry_point=0x10))                       # vm.rs:141-147
    with pytest.raises(z.RuntimeError, match="Decode error"):
        VM(prog([0x7F])).run()


def test_writelog_matches_register_diff():
    """zkir_vm_trace_writelog == the diff of consecutive PRE-state register files (what a Rust recorder would log in
    VMState::write_reg, zkir-runtime/src/state.rs:76-91)."""
    import numpy as np
    from conftest import fib_program_input
    res = VM(fib_program_input(), [40], VMConfig(enable_execution_trace=True)).run()
    rows, wl = res.rows(), res.writelog()
    regs = np.vstack([rows["regs"], rows["final_regs"][None, :]])
    assert np.array_equal(wl["pcs"].astype(np.uint64), rows["pcs"])
    for i in range(res.cycles):
        changed = np.nonzero(regs[i + 1] != regs[i])[0]
        if len(changed) == 0:
            assert wl["wlog"][i] == 0
        else:
            assert len(changed) == 1
            k = int(changed[0])
            assert int(wl["wlog"][i]) == (k << 56) | int(regs[i + 1][k])


def test_poseidon2_witness_records(oracle):
    """Poseidon2Witness (zkir-spec/src/trace.rs:287-304: input_state, output_state, timestamp per SYS_POSEIDON2 call) of a traced run; the
    permutation is the spec's (docs/PROVER_SPEC.md section 2), checked against the oracle's."""
    from zkir_b200.workloads import pos2_program, pos2_cycles
    res = z.VM(pos2_program(), [5], z.VMConfig(enable_execution_trace=True, enable_poseidon2_syscall=True)).run()
    assert res.cycles == pos2_cycles(5)
    ts, ins, outs = res.poseidon2_witness
    assert ts.tolist() == [7 + 16 * k for k in range(5)] and ins.shape == outs.shape == (5, 16)   # cycle of each ecall: 6 set-up cycles, 16 per iteration
    assert np.array_equal(oracle.poseidon2(ins), outs)
    assert ins[0].tolist() == [0] * 16 and np.array_equal(ins[1:], outs[:-1])     # the loop permutes the state in place
    # a run without the syscall has no records; an untraced run records none either
    assert z.run(z.assemble("ebreak")).poseidon2_witness[1].shape == (0, 16)
