#!/usr/bin/env python3
"""Single source of truth for the zkir-b200 AIR (v2: core instruction set + LogUp lookups).

The reference defines no columns and no constraints; the hints it gives are followed here: rows hold
the PRE-state (zkir-runtime/src/vm.rs:234-253,302-312), values are 2x20-bit limbs (zkir-spec/src/value.rs:522-538,592-601)
range-checked as 10-bit chunks against a 1024-entry table (zkir-runtime/src/range_check.rs:170-192,
zkir-spec/src/config.rs:76-80), r0 is hard-wired to zero (zkir-runtime/src/state.rs:76-91).  Transition semantics follow
zkir-runtime/src/execute.rs: ADD :43-63, SUB :65-78, ADDI :185-197, SLTU/SGEU/SEQ/SNE :330-420, CMOV/CMOVZ/CMOVNZ :422-470,
BEQ/BNE/BLTU/BGEU :578-637, JAL :639-647, JALR :649-659, ECALL/EBREAK :661-673, and zkir-runtime/src/syscall.rs:94-149
(EXIT/READ/WRITE/POSEIDON2).

Three column groups (docs/PROVER_SPEC.md section 3):
  main   88 base columns, committed first;
  aux    16 base columns = 4 ext4 values (three LogUp helper sums and the running sum), committed after the lookup
         challenges z, theta are drawn;
  public  4 base columns the verifier evaluates itself (range table 0..1023 and the decoded program ROM), never committed.

This script emits the same constraint list three times, as straight-line code over an abstract context type `C`
(C::F base, C::X ext4; c.L/N main local/next, c.A/AN aux local/next, c.P public column, c.PV public value, c.K constant,
c.z(), c.th(k) lookup challenges, c.xk/xf/x4 ext constructors, c.emit / c.emit_x):
  * zkir_b200/csrc/air_generated.h   -- CUDA quotient kernel (Montgomery u32), CUDA aux-column generator, host verifier (ext4 at zeta)
  * oracle/air_generated.h           -- CPU oracle (canonical u64 arithmetic)
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
# opcode selectors.  `neg` is the polarity of the families that come in pairs whose opcode numbers differ by one
# (BEQ/BNE, SLTU/SGEU, SEQ/SNE, BLTU/BGEU, CMOV/CMOVZ); an ECALL row is is_exit + is_read + is_write + is_pos2;
# a padding row is s_pad = 1 - (all of them): the last member of every "exactly one" group is an expression, not a column
SEL_OPCODE = [("s_add", 0x00), ("s_sub", 0x01), ("s_addi", 0x08), ("s_beq", 0x40), ("s_jal", 0x48), ("s_sltu", 0x20), ("s_seq", 0x24),
              ("s_bltu", 0x44), ("s_cmov", 0x26), ("s_cmovnz", 0x28), ("s_jalr", 0x49), ("s_ebreak", 0x51)]
SEL_NAMES = [n for n, _ in SEL_OPCODE]
S = {n: col(n) for n in SEL_NAMES}
NEG = col("neg")
IS_EXIT, IS_READ, IS_WRITE, IS_POS2 = col("is_exit"), col("is_read"), col("is_write"), col("is_pos2")
ECALL_OPCODE = 0x50
# register indices: index = 4*h + l as a product of two 4-way one-hots; entry 3 of each is 1 - (entries 0..2), so an operand
# costs 6 columns.  rd additionally carries rdw[h] = rd_h[h] * (write enable) -- all four, because the write enable of a
# conditional move is itself a product -- so that the write-back selector rdw[h]*rd_l[l] is degree 2.
RD_H = [col(f"rd_h{i}") for i in range(3)]
RD_L = [col(f"rd_l{i}") for i in range(3)]
RDW = [col(f"rdw{i}") for i in range(4)]
RS1_H = [col(f"rs1_h{i}") for i in range(3)]
RS1_L = [col(f"rs1_l{i}") for i in range(3)]
RS2_H = [col(f"rs2_h{i}") for i in range(3)]
RS2_L = [col(f"rs2_l{i}") for i in range(3)]
A_LO, A_HI, B_LO, B_HI = (col(n) for n in ["a_lo", "a_hi", "b_lo", "b_hi"])
V_LO, V_HI = col("v_lo"), col("v_hi")          # the value written to rd (if the row writes)
# Four 10-bit chunks, looked up in the range table on the rows listed under `rc_on` below: they hold the limbs of v (ADD, SUB,
# ADDI, JAL, READ), of a - b mod 2^40 (SLTU/SGEU/BLTU/BGEU), or v plus a_hi (JALR).  On the other rows the lookups are off and
# the cells carry helper values: ch0/ch1 = inverses of the is-zero gadget, ch2/ch3 = its results (EQ family, CMOV family).
CH = [col(f"ch{i}") for i in range(4)]
CARRY0, CARRY1 = col("carry0"), col("carry1")  # carry / borrow chain; "limb differs" flags of the is-zero gadget
TAKEN = col("taken")                           # branch taken; JALR: bit 0 of rs1 + imm (execute.rs:655 clears it)
M_RNG, M_ROM = col("m_rng"), col("m_rom")      # LogUp multiplicities of the range table row / the ROM row at this trace row
WIDTH = len(COLS)
assert WIDTH == 88   # 11 sponge absorptions per Merkle leaf (rate 8)

# aux columns: ext4 values h0, h1, h2 (helper sums of two fractions each) and phi (running sum), 4 base columns each
AUX_NAMES = ["h0", "h1", "h2", "phi"]
AUX_WIDTH = 4 * len(AUX_NAMES)
# public columns (evaluated by the verifier): range table, ROM pc, ROM decoded word, ROM immediate
PUB_NAMES = ["p_t", "p_pc", "p_dec", "p_imm"]
P_T, P_PC, P_DEC, P_IMM = range(4)
PUB_WIDTH = len(PUB_NAMES)
RANGE_BITS = 10

PV_NAMES = ["entry_pc", "num_cycles", "exit_lo", "exit_hi", "halted"]
NUM_PUBLIC = len(PV_NAMES)


# ----------------------------------------------------------------------------- expression IR
class E:
    """Expression node; .code is C++ over the context `c`; .t is 'F' (base field) or 'X' (ext4)."""

    def __init__(self, code, atom=False, t="F"):
        self.code = code
        self.atom = atom
        self.t = t

    def p(self):
        return self.code if self.atom else f"({self.code})"

    def _bin(self, o, op):
        o = lift(o)
        a, b = self, o
        if a.t != b.t:   # promote the base operand of + / - ; for * put the ext operand on the left (X * F is a scaling)
            if op == "*":
                if a.t == "F":
                    a, b = b, a
                return E(f"{a.p()} * {b.p()}", t="X")
            if a.t == "F":
                a = E(f"c.xf({a.code})", atom=True, t="X")
            else:
                b = E(f"c.xf({b.code})", atom=True, t="X")
        return E(f"{a.p()} {op} {b.p()}", t=a.t)

    def __add__(self, o):
        return self._bin(o, "+")

    __radd__ = lambda self, o: lift(o)._bin(self, "+")

    def __sub__(self, o):
        return self._bin(o, "-")

    def __rsub__(self, o):
        return lift(o)._bin(self, "-")

    def __mul__(self, o):
        return self._bin(o, "*")

    __rmul__ = lambda self, o: lift(o)._bin(self, "*")


def lift(x):
    if isinstance(x, E):
        return x
    return E(f"c.K({int(x) % P}u)", atom=True)


class Gen:
    def __init__(self, pre=""):
        self.lines = []
        self.loaded = {}
        self.ntmp = 0
        self.idx = 0
        self.pre = pre

    def L(self, i):
        return self._ld("l", "L", i, COLS[i])

    def N(self, i):
        return self._ld("n", "N", i, COLS[i])

    def A(self, i):
        return self._ld("a", "A", i, f"{AUX_NAMES[i // 4]}.{i % 4}")

    def AN(self, i):
        return self._ld("an", "AN", i, f"next {AUX_NAMES[i // 4]}.{i % 4}")

    def Pc(self, i):
        return self._ld("p", "P", i, PUB_NAMES[i])

    def _ld(self, pre, fn, i, note):
        key = (pre, i)
        if key not in self.loaded:
            self.lines.append(f"  const F {self.pre}{pre}{i} = c.{fn}({i});  // {note}")
            self.loaded[key] = E(f"{self.pre}{pre}{i}", atom=True)
        return self.loaded[key]

    def PV(self, i):
        return E(f"c.PV({i})", atom=True)

    def tmp(self, e, note=""):
        e = lift(e)
        name = f"{self.pre}{'x' if e.t == 'X' else 't'}{self.ntmp}"
        self.ntmp += 1
        self.lines.append(f"  const {e.t} {name} = {e.code};" + (f"  // {note}" if note else ""))
        return E(name, atom=True, t=e.t)

    def emit(self, e, note):
        e = lift(e)
        fn = "emit_x" if e.t == "X" else "emit"
        self.lines.append(f"  c.{fn}({self.idx}, {e.code});  // {note}")
        self.idx += 1


def sum_e(xs):
    xs = list(xs)
    acc = xs[0]
    for x in xs[1:]:
        acc = acc + x
    return acc


TWO10, TWO20 = 1 << 10, 1 << 20


def shared(g):
    """Row-local expressions used both by the constraints and by the lookup fractions."""
    L = g.L
    s = {n: L(S[n]) for n in SEL_NAMES}
    s_ecall = g.tmp(L(IS_EXIT) + L(IS_READ) + L(IS_WRITE) + L(IS_POS2), "ecall row (syscall.rs:94-149)")
    s_pad = g.tmp(1 - sum_e(s.values()) - s_ecall, "padding row: no other selector set")
    live = g.tmp(1 - s_pad, "live row")

    def idx(H, Lo):
        """register index 4h + l as a LINEAR expression of the one-hot columns (entry 3 is 1 - the others)"""
        h = [L(i) for i in H]
        l = [L(i) for i in Lo]
        hv = 3 - 3 * h[0] - 2 * h[1] - h[2]   # 0*h0 + 1*h1 + 2*h2 + 3*(1 - h0 - h1 - h2)
        lv = 3 - 3 * l[0] - 2 * l[1] - l[2]
        return 4 * hv + lv
    rc_on = g.tmp(s["s_add"] + s["s_addi"] + s["s_sub"] + L(IS_READ) + s["s_jal"] + s["s_jalr"] + s["s_sltu"] + s["s_bltu"], "rows whose chunks are range-checked")
    opcode = g.tmp(sum_e(op * s[n] for n, op in SEL_OPCODE if op) + ECALL_OPCODE * s_ecall + L(NEG), "opcode number (opcode.rs:24-144)")
    rd_eff = g.tmp(idx(RD_H, RD_L) - 10 * (L(IS_READ) + L(IS_POS2)), "rd field of the word (READ / POSEIDON2 write r10 without an rd field)")
    dec = g.tmp(opcode + 128 * rd_eff + 2048 * idx(RS1_H, RS1_L) + 32768 * idx(RS2_H, RS2_L), "decoded word: opcode | rd << 7 | rs1 << 11 | rs2 << 15")
    imm_f = g.tmp(L(IMM_LO) - TWO20 * L(IMM_SIGN), "signed immediate as a field element (execute.rs:187)")
    return s, s_ecall, s_pad, live, rc_on, dec, imm_f


def fractions(g, sh):
    """The 8 LogUp fractions of a row as (numerator F, denominator X).  Bus 1 = 10-bit range table, bus 2 = program ROM, bus 3 = public I/O.
    Fingerprints: range `1 + theta*value`; ROM `2 + theta*pc + theta^2*dec + theta^3*imm`; I/O `3 + theta*clk + theta^2*kind + theta^3*lo + theta^4*hi`."""
    s, s_ecall, s_pad, live, rc_on, dec, imm_f = sh
    L = g.L
    z = E("c.z()", atom=True, t="X")
    th = [None] + [E(f"c.th({k})", atom=True, t="X") for k in (1, 2, 3, 4)]
    out = []
    for j in range(4):
        out.append((rc_on, g.tmp(z - (th[1] * L(CH[j]) + 1), f"range lookup of ch{j}")))
    out.append((g.tmp(0 - L(M_RNG)), g.tmp(z - (th[1] * g.Pc(P_T) + 1), "range table row")))
    out.append((live, g.tmp(z - (th[1] * L(PC) + th[2] * dec + th[3] * imm_f + 2), "ROM lookup of (pc, decoded word, imm)")))
    out.append((g.tmp(0 - L(M_ROM)), g.tmp(z - (th[1] * g.Pc(P_PC) + th[2] * g.Pc(P_DEC) + th[3] * g.Pc(P_IMM) + 2), "ROM table row")))
    # bus 3 = the public I/O transcript: every READ / WRITE row sends (clk, kind, value); the table side is the public list of events,
    # summed by the verifier itself (c.sio() in the closing constraint).  On a WRITE row v holds the written word (r11), on a READ row the tape value.
    out.append((g.tmp(L(IS_READ) + L(IS_WRITE), "I/O row"),
                g.tmp(z - (th[1] * L(CLK) + th[2] * L(IS_WRITE) + th[3] * L(V_LO) + th[4] * L(V_HI) + 3), "I/O event (clk, kind, value)")))
    return out


# helper k sums the fractions FRAC_PAIRS[k]; the two fractions FRAC_PHI are added by the running-sum transition itself
FRAC_PAIRS = [(0, 1), (2, 3), (4, 6)]
FRAC_PHI = (5, 7)
NUM_FRACTIONS = 8


def build():
    g = Gen()
    L, N = g.L, g.N
    first, last, trans = E("c.is_first", True), E("c.is_last", True), E("c.is_trans", True)
    sh = shared(g)
    s, s_ecall, s_pad, live, rc_on, dec, imm_f = sh
    neg = L(NEG)

    def onehot(grp, note):
        """4-way one-hot from 3 columns; the derived entry makes the sum 1 by construction."""
        xs = [L(i) for i in grp]
        return xs + [g.tmp(1 - sum_e(xs), note + "[3]")]
    rd_h, rd_l = onehot(RD_H, "rd.h"), onehot(RD_L, "rd.l")
    rs1_h, rs1_l = onehot(RS1_H, "rs1.h"), onehot(RS1_L, "rs1.l")
    rs2_h, rs2_l = onehot(RS2_H, "rs2.h"), onehot(RS2_L, "rs2.l")

    # --- booleans (derived entries included: with the sums fixed to 1 this makes every group exactly-one-hot)
    for b in [S[n] for n in SEL_NAMES] + [NEG, CARRY0, CARRY1, IMM_SIGN, TAKEN, IS_EXIT, IS_READ, IS_WRITE, IS_POS2]:
        x = L(b)
        g.emit(x * (x - 1), f"bool {COLS[b]}")
    g.emit(s_pad * (s_pad - 1), "bool s_pad (derived): exactly one opcode selector")
    for name, grp in (("rd.h", rd_h), ("rd.l", rd_l), ("rs1.h", rs1_h), ("rs1.l", rs1_l), ("rs2.h", rs2_h), ("rs2.l", rs2_l)):
        for k, x in enumerate(grp):
            g.emit(x * (x - 1), f"bool {name}[{k}]")
    g.emit(neg * (1 - s["s_beq"] - s["s_sltu"] - s["s_seq"] - s["s_bltu"] - s["s_cmov"]), "polarity only on the paired families")

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
    a_lo, a_hi, b_lo, b_hi, v_lo, v_hi = (L(x) for x in (A_LO, A_HI, B_LO, B_HI, V_LO, V_HI))
    ch = [L(x) for x in CH]
    g.emit(a_lo - rs1_lo, "a.lo = reg[rs1].lo")
    g.emit(a_hi - rs1_hi, "a.hi = reg[rs1].hi")
    # ADDI has no rs2 field: the ROM lookup binds its rs2 index to 0 = r0, so b = reg[rs2] + addi * imm stays degree 3
    g.emit(b_lo - rs2_lo - s["s_addi"] * L(IMM_LO), "b.lo = reg[rs2].lo + addi * imm.lo")
    imm_hi = g.tmp((TWO20 - 1) * L(IMM_SIGN), "imm.hi = sign-extension limb")
    g.emit(b_hi - rs2_hi - s["s_addi"] * imm_hi, "b.hi = reg[rs2].hi + addi * imm.hi")
    # --- range-checked pair: rows in rc_on look ch0..ch3 up in the 10-bit table (range_check.rs:175-192)
    rc_lo = g.tmp(ch[0] + TWO10 * ch[1], "range-checked low limb")
    rc_hi = g.tmp(ch[2] + TWO10 * ch[3], "range-checked high limb")
    k0, k1 = L(CARRY0), L(CARRY1)
    # --- ALU (value.rs:620-631 wrap mod 2^40: carry1 is discarded)
    addlike = g.tmp(s["s_add"] + s["s_addi"], "add-like")
    g.emit(addlike * (a_lo + b_lo - v_lo - TWO20 * k0), "add lo limb")
    g.emit(addlike * (a_hi + b_hi + k0 - v_hi - TWO20 * k1), "add hi limb")
    g.emit(s["s_sub"] * (a_lo - b_lo - v_lo + TWO20 * k0), "sub lo limb (carry0 = borrow)")
    g.emit(s["s_sub"] * (a_hi - b_hi - k0 - v_hi + TWO20 * k1), "sub hi limb")
    vrc = g.tmp(addlike + s["s_sub"] + L(IS_READ) + s["s_jal"], "rows whose written value is the range-checked pair")
    g.emit(vrc * (v_lo - rc_lo), "v.lo is range-checked")
    g.emit(vrc * (v_hi - rc_hi), "v.hi is range-checked")
    # --- unsigned compare: the chunks hold a - b mod 2^40, carry1 = final borrow = (a < b)   (execute.rs:330-360, 610-637)
    cmpu = g.tmp(s["s_sltu"] + s["s_bltu"], "unsigned compare row")
    g.emit(cmpu * (a_lo - b_lo - rc_lo + TWO20 * k0), "cmp lo limb")
    g.emit(cmpu * (a_hi - b_hi - k0 - rc_hi + TWO20 * k1), "cmp hi limb")
    lt_x = g.tmp(k1 + neg - 2 * k1 * neg, "(a < b) xor polarity")
    g.emit(s["s_sltu"] * (v_lo - lt_x), "sltu / sgeu result")
    # --- is-zero gadget: EQ family on a - b, CMOV family on b.  carry0/1 = "limb differs", ch0/ch1 = inverses
    eqf = g.tmp(s["s_beq"] + s["s_seq"], "equality row")
    cm = g.tmp(s["s_cmov"] + s["s_cmovnz"], "conditional-move row")
    zf = g.tmp(eqf + cm, "is-zero gadget active")
    x_lo = g.tmp(eqf * (a_lo - b_lo) + cm * b_lo, "gadget input lo")
    x_hi = g.tmp(eqf * (a_hi - b_hi) + cm * b_hi, "gadget input hi")
    g.emit(zf * k0 - x_lo * ch[0], "nz.lo = x.lo * inv.lo")
    g.emit(x_lo * (1 - k0), "x.lo != 0 -> nz.lo = 1")
    g.emit(zf * k1 - x_hi * ch[1], "nz.hi = x.hi * inv.hi")
    g.emit(x_hi * (1 - k1), "x.hi != 0 -> nz.hi = 1")
    nz_cell = g.tmp(eqf * ch[2] + cm * ch[3], "cell that holds nz = nz.lo or nz.hi")
    g.emit(zf * (k0 + k1 - k0 * k1) - nz_cell, "nz = nz.lo or nz.hi")
    ne = ch[2]
    eq_x = g.tmp(1 - ne + neg * (2 * ne - 1), "(a == b) xor polarity")
    g.emit(s["s_seq"] * (v_lo - eq_x), "seq / sne result")
    g.emit((s["s_sltu"] + s["s_seq"]) * v_hi, "set results are 0 / 1")
    # conditional move (execute.rs:422-470): ch3 = (b != 0), ch2 = move flag, v = a
    mv = ch[2]
    g.emit(s["s_cmovnz"] * (mv - ch[3]) + s["s_cmov"] * (mv - ch[3] - neg + 2 * neg * ch[3]), "move flag: cmov/cmovnz b != 0, cmovz b == 0")
    g.emit(cm * (v_lo - a_lo), "cmov value lo")
    g.emit(cm * (v_hi - a_hi), "cmov value hi")
    # --- jumps: link = pc + 4 (execute.rs:639-659)
    g.emit(s["s_jal"] * (v_lo + TWO20 * v_hi - L(PC) - 4), "jal link = pc + 4")
    # JALR: the chunks check the link limbs (v.hi < 2^10: pc < 2^30) and a.hi < 2^10, so that a + imm does not wrap in the field
    g.emit(s["s_jalr"] * (v_lo - rc_lo), "jalr link lo is range-checked")
    g.emit(s["s_jalr"] * (v_hi - ch[2]), "jalr link hi < 2^10")
    g.emit(s["s_jalr"] * (a_hi - ch[3]), "jalr: target base < 2^30")
    g.emit(s["s_jalr"] * (v_lo + TWO20 * v_hi - L(PC) - 4), "jalr link = pc + 4")
    # --- syscalls (syscall.rs:94-149)
    g.emit((L(IS_READ) + L(IS_POS2)) * (rd_h[2] * rd_l[2] - 1), "read / poseidon2 write r10 (syscall.rs:104-109,140-149)")
    g.emit(L(IS_WRITE) * (v_lo - L(REG_LO[11])), "write: v = the written word r11 (lo), sent to the I/O bus (syscall.rs:110-119)")
    g.emit(L(IS_WRITE) * (v_hi - L(REG_HI[11])), "write: v = r11 (hi)")
    g.emit(L(IS_POS2) * v_lo, "poseidon2 returns 0 (lo)")
    g.emit(L(IS_POS2) * v_hi, "poseidon2 returns 0 (hi)")
    # --- register write-back, pre-state rows: next.r[i] = (rd == i && w) ? v : r[i]
    w = g.tmp(addlike + s["s_sub"] + s["s_jal"] + s["s_jalr"] + s["s_sltu"] + s["s_seq"] + L(IS_READ) + L(IS_POS2) + cm * mv, "write enable")
    rdw = [L(RDW[h]) for h in range(4)]
    for h in range(4):
        g.emit(rdw[h] - rd_h[h] * w, f"rdw{h} = rd.h{h} * write enable")
    for i in range(1, 16):
        wi = g.tmp(rdw[i >> 2] * rd_l[i & 3])
        g.emit(trans * (N(REG_LO[i]) - L(REG_LO[i]) - wi * (v_lo - L(REG_LO[i]))), f"write-back r{i}.lo")
        g.emit(trans * (N(REG_HI[i]) - L(REG_HI[i]) - wi * (v_hi - L(REG_HI[i]))), f"write-back r{i}.hi")
    # --- branches
    taken = L(TAKEN)
    g.emit(s["s_beq"] * (taken - eq_x), "beq / bne taken")
    g.emit(s["s_bltu"] * (taken - lt_x), "bltu / bgeu taken")
    br = g.tmp(s["s_beq"] + s["s_bltu"], "branch row")
    g.emit((1 - br - s["s_jalr"]) * taken, "taken only on branch rows (jalr: bit 0 of the target)")
    # --- pc / clk / padding / halting
    # EBREAK leaves the pc where it is (execute.rs:667-673: next_pc = pc), every other live row advances by 4 unless it jumps
    g.emit(trans * (N(PC) - L(PC) - 4 * (live - s["s_ebreak"]) - (br * taken + s["s_jal"]) * (imm_f - 4)
                    - s["s_jalr"] * (a_lo + TWO20 * a_hi + imm_f - taken - L(PC) - 4)), "next pc (jalr: (rs1 + imm) & ~1)")
    g.emit(trans * (N(CLK) - L(CLK) - live), "clk counts live rows")
    # the last row is always a padding row (the packer keeps at least one): the closing constraints below may then ignore its own fractions
    g.emit(last * live, "the last row is a padding row")
    g.emit(last * (L(CLK) - g.PV(1)), "last row: clk = num_cycles")
    n_live = g.tmp(sum_e(N(S[n]) for n in SEL_NAMES) + N(IS_EXIT) + N(IS_READ) + N(IS_WRITE) + N(IS_POS2), "1 - next.s_pad")
    hlt = g.tmp(L(IS_EXIT) + s["s_ebreak"], "halting row (syscall.rs:98-103, execute.rs:667-673)")
    g.emit(trans * (s_pad * n_live), "padding is sticky")
    g.emit(trans * (hlt * n_live), "a halting row is followed by padding")
    g.emit(trans * (g.PV(4) * (live * (1 - n_live) * (1 - hlt))), "halted = 1: padding starts only after a halting row")
    # --- ecall decode (syscall.rs:18-24,94-149): number in r10
    g.emit(L(IS_EXIT) * L(REG_LO[10]), "exit: r10 = 0")
    g.emit(L(IS_READ) * (L(REG_LO[10]) - 1), "read: r10 = 1")
    g.emit(L(IS_WRITE) * (L(REG_LO[10]) - 2), "write: r10 = 2")
    g.emit(L(IS_POS2) * (L(REG_LO[10]) - 4), "poseidon2: r10 = 4")
    g.emit(s_ecall * L(REG_HI[10]), "ecall: r10.hi = 0")
    g.emit(L(IS_EXIT) * (L(REG_LO[11]) - g.PV(2)), "exit code lo (public)")
    g.emit(L(IS_EXIT) * (L(REG_HI[11]) - g.PV(3)), "exit code hi (public)")
    g.emit(s["s_ebreak"] * g.PV(2), "ebreak: no exit code (lo)")
    g.emit(s["s_ebreak"] * g.PV(3), "ebreak: no exit code (hi)")
    # --- first row (vm.rs:149,177-181; state.rs:55-71)
    g.emit(first * L(CLK), "clk0 = 0")
    g.emit(first * (L(PC) - g.PV(0)), "pc0 = entry point")
    for i in range(1, 16):
        g.emit(first * L(REG_LO[i]), f"r{i}.lo starts 0")
        g.emit(first * L(REG_HI[i]), f"r{i}.hi starts 0")
    # --- LogUp: helper k = n_i/d_i + n_j/d_j; phi' = phi + h0 + h1 + h2 + n_5/d_5; phi_first = 0; the last row closes the sum to 0
    fr = fractions(g, sh)

    def xaux(k, nxt=False):
        ld = g.AN if nxt else g.A
        cs = [ld(4 * k + j) for j in range(4)]
        return g.tmp(E(f"c.x4({cs[0].code}, {cs[1].code}, {cs[2].code}, {cs[3].code})", atom=True, t="X"), ("next " if nxt else "") + AUX_NAMES[k])
    h = [xaux(k) for k in range(3)]
    phi, phi_n = xaux(3), xaux(3, True)
    for k, (i, j) in enumerate(FRAC_PAIRS):
        (ni, di), (nj, dj) = fr[i], fr[j]
        g.emit(h[k] * di * dj - di * nj - dj * ni, f"helper {k} = fraction {i} + fraction {j}")
    (n5, d5), (n7, d7) = fr[FRAC_PHI[0]], fr[FRAC_PHI[1]]
    hs = g.tmp(h[0] + h[1] + h[2], "h0 + h1 + h2")
    sio = E("c.sio()", atom=True, t="X")
    g.emit(((phi_n - phi - hs) * d5 * d7 - d7 * n5 - d5 * n7) * trans, "running sum transition (adds the ROM lookup and the I/O event itself)")
    g.emit(phi * first, "running sum starts at 0")
    # last row = padding row: its ROM-lookup and I/O numerators are 0 (live = 0), only the table-side helpers count
    g.emit((sio - phi - hs) * last, "range and ROM lookups cancel; what remains is the public I/O transcript's sum")
    return g


def build_fractions():
    g = Gen(pre="f_")
    fr = fractions(g, shared(g))
    for j, (n, d) in enumerate(fr):
        g.lines.append(f"  c.frac({j}, {lift(n).code}, {d.code});")
    return g


def main():
    g = build()
    gf = build_fractions()
    hdr = []
    hdr.append(f"// GENERATED by tools/gen_air.py -- do not edit.  zkir-b200 AIR v2 ({WIDTH} main + {AUX_WIDTH} aux + {PUB_WIDTH} public columns).")
    hdr.append("#pragma once")
    hdr.append(f"#define ZKIR_AIR_WIDTH {WIDTH}")
    hdr.append(f"#define ZKIR_AIR_AUX_WIDTH {AUX_WIDTH}")
    hdr.append(f"#define ZKIR_AIR_PUB_WIDTH {PUB_WIDTH}")
    hdr.append(f"#define ZKIR_AIR_NUM_CONSTRAINTS {g.idx}")
    hdr.append(f"#define ZKIR_AIR_NUM_PUBLIC {NUM_PUBLIC}")
    hdr.append(f"#define ZKIR_AIR_NUM_FRACTIONS {NUM_FRACTIONS}")
    hdr.append(f"#define ZKIR_AIR_RANGE_BITS {RANGE_BITS}")
    hdr.append("#define ZKIR_AIR_MAX_DEGREE 3")
    hdr.append("// fraction j is summed by helper ZKIR_AIR_FRAC_HELPER[j] (3 = added by the running-sum transition itself)")
    helper_of = [3] * NUM_FRACTIONS   # 3 = the running sum itself
    for k, (i, j) in enumerate(FRAC_PAIRS):
        helper_of[i] = helper_of[j] = k
    hdr.append("#define ZKIR_AIR_FRAC_HELPER_INIT {" + ", ".join(str(x) for x in helper_of) + "}")
    hdr.append("#ifndef ZKIR_HD\n#ifdef __CUDACC__\n#define ZKIR_HD __host__ __device__ __forceinline__\n#else\n#define ZKIR_HD inline\n#endif\n#endif")
    hdr.append("// Context contract: C::F (base) and C::X (ext4) with + - *, X * F scaling; c.L(i)/c.N(i) main local/next row, c.A(i)/c.AN(i) aux,")
    hdr.append("// c.P(i) public column, c.PV(i) public value, c.K(u32 canonical constant), c.z()/c.th(k) lookup challenges z, theta^k,")
    hdr.append("// c.xf(F) -> X, c.x4(F,F,F,F) -> X, c.sio() -> X (sum of the public I/O transcript's fractions), c.is_first / c.is_last / c.is_trans selectors,")
    hdr.append("// c.emit(index, F), c.emit_x(index, X).")
    hdr.append("template <class C> ZKIR_HD void zkir_air_eval(C& c) {")
    hdr.append("  typedef typename C::F F;")
    hdr.append("  typedef typename C::X X;")
    hdr.extend(g.lines)
    hdr.append("}")
    hdr.append("// The LogUp fractions of one row, c.frac(j, numerator F, denominator X): used to BUILD the aux columns (local row only).")
    hdr.append("template <class C> ZKIR_HD void zkir_air_fractions(C& c) {")
    hdr.append("  typedef typename C::F F;")
    hdr.append("  typedef typename C::X X;")
    hdr.extend(gf.lines)
    hdr.append("}")
    text = "\n".join(hdr) + "\n"
    for rel in ("zkir_b200/csrc/air_generated.h", "oracle/air_generated.h"):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(text)
    # column map: C header for the packer + python module for tests
    ch = ["// GENERATED by tools/gen_air.py -- column indices of the AIR v2.", "#pragma once"]
    for i, n in enumerate(COLS):
        ch.append(f"#define ZKIR_COL_{n.upper()} {i}")
    ch.append(f"#define ZKIR_COL_COUNT {WIDTH}")
    for i, n in enumerate(PUB_NAMES):
        ch.append(f"#define ZKIR_PUB_{n.upper()} {i}")
    with open(os.path.join(ROOT, "zkir_b200/csrc/air_columns.h"), "w") as f:
        f.write("\n".join(ch) + "\n")
    with open(os.path.join(ROOT, "zkir_b200/air_layout.py"), "w") as f:
        f.write('"""GENERATED by tools/gen_air.py -- column map of the AIR v2."""\n')
        f.write(f"WIDTH = {WIDTH}\nAUX_WIDTH = {AUX_WIDTH}\nPUB_WIDTH = {PUB_WIDTH}\nNUM_CONSTRAINTS = {g.idx}\nNUM_PUBLIC = {NUM_PUBLIC}\n")
        f.write(f"MIN_LOG_N = {RANGE_BITS}\n")
        f.write(f"PUBLIC_NAMES = {PV_NAMES!r}\n")
        f.write("COLUMNS = " + repr(COLS) + "\n")
        f.write("INDEX = {n: i for i, n in enumerate(COLUMNS)}\n")
    print(f"AIR: width={WIDTH} aux={AUX_WIDTH} pub={PUB_WIDTH} constraints={g.idx}")


if __name__ == "__main__":
    main()
