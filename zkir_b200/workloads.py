"""Synthetic workloads of the BASELINE configs, shared by bench.py, the smoke check and the tests.

Programs are written in the v3.4 assembly the reference's own tests use (tests/end_to_end.rs:310-332 for the fibonacci loop,
syscall numbers in R10 per zkir-runtime/src/syscall.rs:82-87)."""
import time

FIB_SRC = """
addi r1, r0, 0
addi r2, r0, 1
addi r3, r0, {n}
addi r4, r0, 2
add r5, r1, r2
add r1, r0, r2
add r2, r0, r5
addi r4, r4, 1
bne r4, r3, -16
add r10, r0, r0
ecall
"""

# BASELINE config 4: a + b from the input tape, 11 cycles
ADD_SRC = ("addi r10, r0, 1\necall\nadd r1, r10, r0\naddi r10, r0, 1\necall\nadd r11, r1, r10\naddi r10, r0, 2\necall\n"
           "addi r10, r0, 0\naddi r11, r0, 0\necall\n")

FIB_N_FULL = 209715          # 5n - 2 = 1_048_573 cycles -> 2^20 rows (BASELINE config 2)


def fib_program(n):
    """BASELINE config 1: the reference's runnable fibonacci with `addi r3,r0,n` (n fits the 17-bit immediate)."""
    from . import assemble
    return assemble(FIB_SRC.format(n=n))


def fib_program_input():
    """BASELINE config 2: n comes from the input tape (17-bit immediates cannot hold 209715; encoder.rs:117)."""
    from . import assemble
    body = FIB_SRC.format(n=0).strip().splitlines()
    return assemble("addi r10, r0, 1\necall\nadd r3, r10, r0\n" + "\n".join(body[:2] + body[3:]))


def add_program():
    from . import assemble
    return assemble(ADD_SRC)


def run_traced(prog, inputs=(), max_cycles=1 << 26):
    from . import VM, VMConfig
    return VM(prog, list(inputs), VMConfig(max_cycles=max_cycles, enable_execution_trace=True)).run()


def fib_trace(n=None, n_input=None, log_n=None):
    """-> (ExecutionResult, columns, public values) of the fibonacci workload (host packer)."""
    if n_input is not None:
        prog, inputs = fib_program_input(), [n_input]
    else:
        prog, inputs = fib_program(n), []
    res = run_traced(prog, inputs)
    cols, pv = res.pack(log_n)
    return res, cols, pv


def timed_fib_trace(n_input):
    """-> (ExecutionResult, cols, pv, interpreter seconds, packer seconds)"""
    prog = fib_program_input()
    t0 = time.perf_counter()
    res = run_traced(prog, [n_input])
    t1 = time.perf_counter()
    cols, pv = res.pack(None)
    return res, cols, pv, t1 - t0, time.perf_counter() - t1


# BASELINE config 3: a loop of SYS_POSEIDON2 over a 16-word state in guest memory, 16 cycles per iteration
# (syscall 4 in R10, state address in R11, output address in R13: syscall.rs:18-24,140-149)
POS2_ITER_CYCLES = 16
POS2_SRC = ("addi r10, r0, 1\necall\nadd r3, r10, r0\n"            # iterations from the input tape
            "addi r11, r0, 0x2000\naddi r13, r0, 0x2000\naddi r4, r0, 0\n"
            "addi r10, r0, 4\necall\naddi r4, r4, 1\n"               # loop body: permute the state in place, count
            + "add r5, r5, r4\n" * 12 +
            "bne r4, r3, -60\n"
            "add r10, r0, r0\nadd r11, r5, r0\necall\n")
POS2_SETUP_CYCLES = 6 + 3
POS2_ITERS_FULL = (1 << 18) - 1     # 16 * (2^18 - 1) + 9 = 4_194_297 cycles -> 2^22 rows


def pos2_program():
    from . import assemble
    return assemble(POS2_SRC)


def pos2_cycles(iters):
    return POS2_ITER_CYCLES * iters + POS2_SETUP_CYCLES


# Full-ISA workload (docs/PROVER_SPEC.md sections 3.7, 3.8): a xorshift-multiply generator loop that exercises the multiplier block (MUL,
# MULH, DIVU, REMU), all three shift kinds, the bitwise table, the signed compares and the memory argument (a 1 KiB table of 8-byte
# slots written and read at data-dependent offsets); 22 cycles per iteration.
MIX_ITER_CYCLES = 22
MIX_SETUP_CYCLES = 8 + 3
MIX_SRC = ("addi r10, r0, 1\necall\nadd r3, r10, r0\n"            # iterations from the input tape
           "addi r1, r0, 12345\naddi r2, r0, 25173\naddi r4, r0, 0\naddi r6, r0, 1000\naddi r9, r0, 0\n"
           "mul r1, r1, r2\n"          # loop body
           "addi r1, r1, 13849\n"
           "srli r5, r1, 7\n"
           "xor r1, r1, r5\n"
           "slli r5, r1, 9\n"
           "xor r1, r1, r5\n"
           "mulh r7, r1, r2\n"
           "or r7, r7, r1\n"
           "andi r7, r7, 1023\n"
           "divu r8, r1, r6\n"
           "remu r8, r8, r6\n"
           "srai r5, r1, 3\n"
           "slt r5, r5, r7\n"
           "add r9, r9, r5\n"
           "andi r12, r1, 1016\n"
           "sd r1, 0x4000(r12)\n"
           "andi r14, r7, 1016\n"
           "ld r13, 0x4000(r14)\n"
           "andi r13, r13, 1\n"
           "add r9, r9, r13\n"
           "addi r4, r4, 1\n"
           "bne r4, r3, -84\n"
           "add r10, r0, r0\nadd r11, r9, r0\necall\n")


def mix_program():
    from . import assemble
    return assemble(MIX_SRC)


def mix_cycles(iters):
    return MIX_ITER_CYCLES * iters + MIX_SETUP_CYCLES


def mix_reference(iters):
    """The loop restated in Python (semantics of zkir-runtime/src/execute.rs: everything wraps at 40 bits, SRAI / SLT are signed at bit 39)."""
    M = (1 << 40) - 1
    sg = lambda v: v - (1 << 40) if v >> 39 else v
    r1, r2, r9, table = 12345, 25173, 0, {}
    for _ in range(iters):
        r1 = (r1 * r2) & M
        r1 = (r1 + 13849) & M
        r1 ^= r1 >> 7
        r1 ^= (r1 << 9) & M
        r7 = ((r1 * r2) >> 40) & M
        r7 = (r7 | r1) & 1023
        r5 = (sg(r1) >> 3) & M
        r9 += int(sg(r5) < sg(r7))
        table[r1 & 1016] = r1
        r9 += table.get(r7 & 1016, 0) & 1
    return r9
