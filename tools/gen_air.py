#!/usr/bin/env python3
"""Single source of truth for the zkir-b200 "core" AIR (v1).

The reference defines no columns and no constraints; the hints it gives are
followed here: rows hold the PRE-state (zkir-runtime/src/vm.rs:234-253,302-312), values are 2x20-bit
limbs (zkir-spec/src/value.rs:522-538,592-601), r0 is hard-wired to zero
(zkir-runtime/src/state.rs:76-91), transition semantics follow zkir-runtime/src/execute.rs
(ADD :43-63, SUB :65-78, ADDI :185-197, BEQ/BNE :578-596, JAL :639-647, ECALL :661-665) and
zkir-runtime/src/syscall.rs:94-119 (EXIT/READ/WRITE).

This script emits the same constraint list three times, as straight-line code over an abstract
context type `C` (fields: C::F, c.L(i), c.N(i), c.PV(i), c.is_first/is_last/is_trans, c.K(u32),
c.emit(idx, expr)):
  * zkir_b200/csrc/air_generated.h       -- instantiated by the CUDA quotient kernel (Montgomery u32)
                                             and by the host verifier (ext4 at zeta)
  * oracle/air_generated.h               -- instantiated by the CPU oracle (canonical u64 arithmetic)
plus the column map (zkir_b200/air_layout.py and a C header) used by the packer.
"""
import os

P = 2013265921
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ----------------------------------------------------------------------------- column layout
COLS = []


def col(name):
    COLS.append(name)
    return len(COLS) - 1


CLK = col("clk")
PC = col("pc")
IMM_LO = col("imm_lo")      # low 20 bits of the sign-extended 40-bit immediate
IMM_SIGN = col("imm_sign")  # its high limb is sign * (2^20 - 1): immediates are 17 / 21 bits (encoder.rs:98-151)
# register file, PRE-state; r0 is hard-wired to zero (state.rs:76-91) and has no columns
REG_LO = [None]
REG_HI = [None]
for i in range(1, 16):
    REG_LO.append(col(f"r{i}_lo"))
    REG_HI.append(col(f"r{i}_hi"))
# opcode selectors; an ECALL row is is_exit + is_read + is_write (no separate column), and a padding row is
# s_pad = 1 - (all other selectors): the last member of every "exactly one" group is a linear expression, not a column
SEL_NAMES = ["s_add", "s_sub", "s_addi", "s_beq", "s_bne", "s_jal"]
S = {n: col(n) for n in SEL_NAMES}
# register indices: index = 4*h + l as a product of two 4-way one-hots; entry 3 of each is 1 - (entries 0..2), so an operand
# costs 6 columns.  rd additionally carries rdw[h] = rd_h[h] * (write enable) for h < 3 (rdw[3] = w - rdw[0] - rdw[1] - rdw[2]),
# so that the write-back selector rdw[h]*rd_l[l] is degree 2.
RD_H = [col(f"rd_h{i}") for i in range(3)]
RD_L = [col(f"rd_l{i}") for i in range(3)]
RDW = [col(f"rdw{i}") for i in range(3)]
RS1_H = [col(f"rs1_h{i}") for i in range(3)]
RS1_L = [col(f"rs1_l{i}") for i in range(3)]
RS2_H = [col(f"rs2_h{i}") for i in range(3)]
RS2_L = [col(f"rs2_l{i}") for i in range(3)]
A_LO, A_HI, B_LO, B_HI, C_LO, C_HI = (col(n) for n in ["a_lo", "a_hi", "b_lo", "b_hi", "c_lo", "c_hi"])
CARRY0, CARRY1 = col("carry0"), col("carry1")
# Branch rows have no result, no carries and no destination register, so their helper values live in the cells the other row
# kinds use for those (every constraint on these cells is gated by the row kind):
#   c_lo / c_hi      = inverses of the limb differences a - b            (ALU / JAL / READ rows: the result limbs)
#   carry0 / carry1  = flags "limb differs"                               (ADD / SUB / ADDI rows: carry / borrow)
#   rd_l[1]          = branch taken (B-type words have no rd field)       (writing rows: a bit of the rd index)
IS_EXIT, IS_READ, IS_WRITE = col("is_exit"), col("is_read"), col("is_write")
WIDTH = len(COLS)
assert WIDTH == 72   # 9 sponge absorptions per Merkle leaf (rate 8)

PV_NAMES = ["entry_pc", "num_cycles", "exit_lo", "exit_hi"]
NUM_PUBLIC = len(PV_NAMES)


# ----------------------------------------------------------------------------- expression IR
class E:
    """Expression node; str(e) is C++ over the context `c`."""

    def __init__(self, code, atom=False):
        self.code = code
        self.atom = atom

    def p(self):
        return self.code if self.atom else f"({self.code})"

    def __add__(self, o):
        return E(f"{self.p()} + {lift(o).p()}")

    __radd__ = lambda self, o: lift(o) + self

    def __sub__(self, o):
        return E(f"{self.p()} - {lift(o).p()}")

    def __rsub__(self, o):
        return lift(o) - self

    def __mul__(self, o):
        return E(f"{self.p()} * {lift(o).p()}")

    __rmul__ = lambda self, o: lift(o) * self


def lift(x):
    if isinstance(x, E):
        return x
    return E(f"c.K({int(x) % P}u)", atom=True)


class Gen:
    def __init__(self):
        self.lines = []
        self.loaded = {}
        self.ntmp = 0
        self.idx = 0

    def L(self, i):
        return self._ld("l", "L", i)

    def N(self, i):
        return self._ld("n", "N", i)

    def _ld(self, pre, fn, i):
        key = (pre, i)
        if key not in self.loaded:
            self.lines.append(f"  const F {pre}{i} = c.{fn}({i});  // {COLS[i]}")
            self.loaded[key] = E(f"{pre}{i}", atom=True)
        return self.loaded[key]

    def PV(self, i):
        return E(f"c.PV({i})", atom=True)

    def tmp(self, e, note=""):
        name = f"t{self.ntmp}"
        self.ntmp += 1
        self.lines.append(f"  const F {name} = {lift(e).code};" + (f"  // {note}" if note else ""))
        return E(name, atom=True)

    def emit(self, e, note):
        self.lines.append(f"  c.emit({self.idx}, {lift(e).code});  // {note}")
        self.idx += 1


def sum_e(xs):
    xs = list(xs)
    acc = xs[0]
    for x in xs[1:]:
        acc = acc + x
    return acc


def build():
    g = Gen()
    L, N = g.L, g.N
    first, last, trans = E("c.is_first", True), E("c.is_last", True), E("c.is_trans", True)
    TWO20 = 1 << 20
    s = {n: L(S[n]) for n in SEL_NAMES}
    s_ecall = g.tmp(L(IS_EXIT) + L(IS_READ) + L(IS_WRITE), "ecall row (syscall.rs:94-119)")
    s["s_pad"] = g.tmp(1 - sum_e(s.values()) - s_ecall, "padding row: no other selector set")

    def onehot(grp, note):
        """4-way one-hot from 3 columns; the derived entry makes the sum 1 by construction."""
        xs = [L(i) for i in grp]
        return xs + [g.tmp(1 - sum_e(xs), note + "[3]")]
    rd_h, rd_l = onehot(RD_H, "rd.h"), onehot(RD_L, "rd.l")
    rs1_h, rs1_l = onehot(RS1_H, "rs1.h"), onehot(RS1_L, "rs1.l")
    rs2_h, rs2_l = onehot(RS2_H, "rs2.h"), onehot(RS2_L, "rs2.l")

    # --- booleans (derived entries included: with the sums fixed to 1 this makes every group exactly-one-hot)
    for b in [S[n] for n in SEL_NAMES] + [CARRY0, CARRY1, IMM_SIGN, IS_EXIT, IS_READ, IS_WRITE]:
        x = L(b)
        g.emit(x * (x - 1), f"bool {COLS[b]}")
    g.emit(s["s_pad"] * (s["s_pad"] - 1), "bool s_pad (derived): exactly one opcode selector")
    for name, grp in (("rd.h", rd_h), ("rd.l", rd_l), ("rs1.h", rs1_h), ("rs1.l", rs1_l), ("rs2.h", rs2_h), ("rs2.l", rs2_l)):
        for k, x in enumerate(grp):
            g.emit(x * (x - 1), f"bool {name}[{k}]")

    # --- operand fetch: reg[4h+l] selected by H[h]*L[l]; r0 contributes nothing (state.rs:76-91)
    def fetch(H, Lo, limb, note):
        terms = []
        for h in range(4):
            inner = [Lo[l] * L(limb[4 * h + l]) for l in range(4) if 4 * h + l != 0]
            terms.append(H[h] * sum_e(inner))
        return g.tmp(sum_e(terms), note)
    rs1_lo = fetch(rs1_h, rs1_l, REG_LO, "rs1.lo")
    rs1_hi = fetch(rs1_h, rs1_l, REG_HI, "rs1.hi")
    rs2_lo = fetch(rs2_h, rs2_l, REG_LO, "rs2.lo")
    rs2_hi = fetch(rs2_h, rs2_l, REG_HI, "rs2.hi")
    a_lo, a_hi, b_lo, b_hi, c_lo, c_hi = (L(x) for x in (A_LO, A_HI, B_LO, B_HI, C_LO, C_HI))
    g.emit(a_lo - rs1_lo, "a.lo = reg[rs1].lo")
    g.emit(a_hi - rs1_hi, "a.hi = reg[rs1].hi")
    # ADDI has no rs2: the converter selects r0 there, which is enforced, so b = reg[rs2] + addi * imm stays degree 3
    g.emit(s["s_addi"] * (1 - rs2_h[0] * rs2_l[0]), "addi: rs2 selector points at r0")
    g.emit(b_lo - rs2_lo - s["s_addi"] * L(IMM_LO), "b.lo = reg[rs2].lo + addi * imm.lo")
    imm_hi = g.tmp((TWO20 - 1) * L(IMM_SIGN), "imm.hi = sign-extension limb")
    g.emit(b_hi - rs2_hi - s["s_addi"] * imm_hi, "b.hi = reg[rs2].hi + addi * imm.hi")
    # --- immediate as a field element: imm_lo + 2^20 imm_hi - sign * 2^40 = imm_lo - 2^20 sign  (execute.rs:187)
    imm_f = g.tmp(L(IMM_LO) - TWO20 * L(IMM_SIGN), "signed immediate")
    # --- ALU (value.rs:620-631 wrap mod 2^40: carry1 is discarded)
    addlike = g.tmp(s["s_add"] + s["s_addi"], "add-like")
    k0, k1 = L(CARRY0), L(CARRY1)
    g.emit(addlike * (a_lo + b_lo - c_lo - TWO20 * k0), "add lo limb")
    g.emit(addlike * (a_hi + b_hi + k0 - c_hi - TWO20 * k1), "add hi limb")
    g.emit(s["s_sub"] * (a_lo - b_lo - c_lo + TWO20 * k0), "sub lo limb (carry0 = borrow)")
    g.emit(s["s_sub"] * (a_hi - b_hi - k0 - c_hi + TWO20 * k1), "sub hi limb")
    g.emit(s["s_jal"] * (c_lo + TWO20 * c_hi - L(PC) - 4), "jal link = pc + 4 (execute.rs:639-647)")
    g.emit(L(IS_READ) * (rd_h[2] * rd_l[2] - 1), "read writes r10 (syscall.rs:104-109); c = the tape value")
    # --- register write-back, pre-state rows: next.r[i] = (rd == i && w) ? c : r[i]
    w = g.tmp(s["s_add"] + s["s_sub"] + s["s_addi"] + s["s_jal"] + L(IS_READ), "write enable")
    rdw = [L(RDW[h]) for h in range(3)]
    for h in range(3):
        g.emit(rdw[h] - rd_h[h] * w, f"rdw{h} = rd.h{h} * write enable")
    rdw.append(g.tmp(w - sum_e(rdw), "rdw3 = rd.h3 * write enable, implied by the three above"))
    for i in range(1, 16):
        wi = g.tmp(rdw[i >> 2] * rd_l[i & 3])
        g.emit(trans * (N(REG_LO[i]) - L(REG_LO[i]) - wi * (c_lo - L(REG_LO[i]))), f"write-back r{i}.lo")
        g.emit(trans * (N(REG_HI[i]) - L(REG_HI[i]) - wi * (c_hi - L(REG_HI[i]))), f"write-back r{i}.hi")
    # --- branches: raw equality of both limbs (execute.rs:578-596)
    d_lo = g.tmp(a_lo - b_lo)
    d_hi = g.tmp(a_hi - b_hi)
    ne_lo, ne_hi, inv_lo, inv_hi, taken = k0, k1, c_lo, c_hi, rd_l[1]   # shared cells, see the column layout
    br = g.tmp(s["s_beq"] + s["s_bne"], "branch row")
    g.emit(br * (ne_lo - d_lo * inv_lo), "branch: ne.lo = d.lo * inv.lo")
    g.emit(br * (d_lo * (1 - ne_lo)), "branch: d.lo != 0 -> ne.lo = 1")
    g.emit(br * (ne_hi - d_hi * inv_hi), "branch: ne.hi = d.hi * inv.hi")
    g.emit(br * (d_hi * (1 - ne_hi)), "branch: d.hi != 0 -> ne.hi = 1")
    ne = g.tmp(ne_lo + ne_hi - ne_lo * ne_hi, "a != b")
    g.emit(s["s_bne"] * (taken - ne) + s["s_beq"] * (taken - 1 + ne), "branch taken (exactly one selector is set on a row)")
    # --- pc / clk / padding
    live = g.tmp(1 - s["s_pad"])
    g.emit(trans * (N(PC) - L(PC) - 4 * live - (br * taken + s["s_jal"]) * (imm_f - 4)), "next pc")
    g.emit(trans * (N(CLK) - L(CLK) - live), "clk counts live rows")
    g.emit(last * (L(CLK) + live - g.PV(1)), "last row: clk (+1 if live) = num_cycles")
    n_live = g.tmp(sum_e(N(S[n]) for n in SEL_NAMES) + N(IS_EXIT) + N(IS_READ) + N(IS_WRITE), "1 - next.s_pad")
    g.emit(trans * (s["s_pad"] * n_live), "padding is sticky")
    g.emit(trans * (L(IS_EXIT) * n_live), "exit is followed by padding")
    # --- ecall decode (syscall.rs:18-24,94-119): number in r10
    g.emit(L(IS_EXIT) * L(REG_LO[10]), "exit: r10 = 0")
    g.emit(L(IS_READ) * (L(REG_LO[10]) - 1), "read: r10 = 1")
    g.emit(L(IS_WRITE) * (L(REG_LO[10]) - 2), "write: r10 = 2")
    g.emit(s_ecall * L(REG_HI[10]), "ecall: r10.hi = 0")
    g.emit(L(IS_EXIT) * (L(REG_LO[11]) - g.PV(2)), "exit code lo (public)")
    g.emit(L(IS_EXIT) * (L(REG_HI[11]) - g.PV(3)), "exit code hi (public)")
    # --- first row (vm.rs:149,177-181; state.rs:55-71)
    g.emit(first * L(CLK), "clk0 = 0")
    g.emit(first * (L(PC) - g.PV(0)), "pc0 = entry point")
    for i in range(1, 16):
        g.emit(first * L(REG_LO[i]), f"r{i}.lo starts 0")
        g.emit(first * L(REG_HI[i]), f"r{i}.hi starts 0")
    return g


def main():
    g = build()
    hdr = []
    hdr.append(f"// GENERATED by tools/gen_air.py -- do not edit.  zkir-b200 core AIR v1 ({WIDTH} columns).")
    hdr.append("#pragma once")
    hdr.append(f"#define ZKIR_AIR_WIDTH {WIDTH}")
    hdr.append(f"#define ZKIR_AIR_NUM_CONSTRAINTS {g.idx}")
    hdr.append(f"#define ZKIR_AIR_NUM_PUBLIC {NUM_PUBLIC}")
    hdr.append("#define ZKIR_AIR_MAX_DEGREE 3")
    hdr.append("#ifndef ZKIR_HD\n#ifdef __CUDACC__\n#define ZKIR_HD __host__ __device__ __forceinline__\n#else\n#define ZKIR_HD inline\n#endif\n#endif")
    hdr.append("// Context contract: typename C::F with + - *; c.L(i) local row, c.N(i) next row, c.PV(i) public value,")
    hdr.append("// c.K(u32 canonical constant), c.is_first / c.is_last / c.is_trans selectors, c.emit(index, value).")
    hdr.append("template <class C> ZKIR_HD void zkir_air_eval(C& c) {")
    hdr.append("  typedef typename C::F F;")
    hdr.extend(g.lines)
    hdr.append("}")
    text = "\n".join(hdr) + "\n"
    for rel in ("zkir_b200/csrc/air_generated.h", "oracle/air_generated.h"):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(text)
    # column map: C header for the packer + python module for tests
    ch = ["// GENERATED by tools/gen_air.py -- column indices of the core AIR v1.", "#pragma once"]
    for i, n in enumerate(COLS):
        ch.append(f"#define ZKIR_COL_{n.upper()} {i}")
    ch.append(f"#define ZKIR_COL_COUNT {WIDTH}")
    with open(os.path.join(ROOT, "zkir_b200/csrc/air_columns.h"), "w") as f:
        f.write("\n".join(ch) + "\n")
    with open(os.path.join(ROOT, "zkir_b200/air_layout.py"), "w") as f:
        f.write('"""GENERATED by tools/gen_air.py -- column map of the core AIR v1."""\n')
        f.write(f"WIDTH = {WIDTH}\nNUM_CONSTRAINTS = {g.idx}\nNUM_PUBLIC = {NUM_PUBLIC}\n")
        f.write(f"PUBLIC_NAMES = {PV_NAMES!r}\n")
        f.write("COLUMNS = " + repr(COLS) + "\n")
        f.write("INDEX = {n: i for i, n in enumerate(COLUMNS)}\n")
    print(f"AIR: width={WIDTH} constraints={g.idx}")


if __name__ == "__main__":
    main()
