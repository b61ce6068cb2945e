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
# Two AIR profiles are generated from this one description (docs/PROVER_SPEC.md section 3.7):
#   core  the 18 opcodes of the headline workloads, 88 + 16 + 4 columns (this file's round-2 layout, unchanged);
#   full  every opcode of zkir-spec/src/opcode.rs:24-144: the core columns plus a 40x40-bit multiplier block (MUL MULH DIVU REMU
#         DIV REM and all six shifts), sign extraction (SLT SGE BLT BGE SRA SRAI), a 5-bit x 5-bit AND table (AND OR XOR ANDI ORI
#         XORI) and an offline memory-checking argument over aligned 8-byte words (LB LBU LH LHU LW LD SB SH SW SD).
# The profile of a proof is its trace width (zkir_params.width): a program that stays inside the core opcodes is proven with the
# narrow table; the verifier needs no flag because an opcode outside a profile has no selector there and matches no ROM row.
FULL = os.environ.get("ZKIR_AIR_PROFILE", "core") == "full"

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
# ----- full profile: appended after the core columns, so every ZKIR_COL_* index of the core layout is valid in both
FULL_SEL_OPCODE = [("s_mul", 0x02), ("s_divu", 0x04), ("s_div", 0x06),                       # + neg: MULH, REMU, REM
                   ("s_and", 0x10), ("s_or", 0x11), ("s_xor", 0x12), ("s_andi", 0x13), ("s_ori", 0x14), ("s_xori", 0x15),
                   ("s_sll", 0x18), ("s_srl", 0x19), ("s_sra", 0x1A), ("s_slli", 0x1B), ("s_srli", 0x1C), ("s_srai", 0x1D),
                   ("s_slt", 0x22), ("s_blt", 0x42),                                          # + neg: SGE, BGE
                   ("s_lb", 0x30), ("s_lbu", 0x31), ("s_lh", 0x32), ("s_lhu", 0x33), ("s_lw", 0x34), ("s_ld", 0x35),
                   ("s_sb", 0x38), ("s_sh", 0x39), ("s_sw", 0x3A), ("s_sd", 0x3B)]
if FULL:
    SEL_OPCODE = SEL_OPCODE + FULL_SEL_OPCODE
    SEL_NAMES = [n for n, _ in SEL_OPCODE]
    for n, _ in FULL_SEL_OPCODE:
        S[n] = col(n)
    # multiplier block: X * Y + R = P over the integers, 10-bit chunks (x, y, r: 4 each, p: 8), carries of the five low columns
    # as two chunks each.  X = rs1 (MUL, shifts, compares, bitwise) or the quotient (DIV family); Y = rs2 or 2^s / 2^(40-s) from
    # the power table (shifts); R = the remainder (DIV family) or rs2 (shifts: its low chunk holds the shift amount)
    XC = [col(f"x{i}") for i in range(4)]
    YC = [col(f"y{i}") for i in range(4)]
    RC = [col(f"r{i}") for i in range(4)]
    PC8 = [col(f"p{i}") for i in range(8)]
    KLO = [col(f"k{i}_lo") for i in range(5)]
    KHI = [col(f"k{i}_hi") for i in range(5)]
    SA, SB = col("sign_a"), col("sign_b")          # bit 39 of X / Y
    SXOR, LTS = col("sign_xor"), col("lt_signed")  # sign_a xor sign_b; (a <u b) xor sign_xor = (a <s b)
    SHAMT, SHW = col("shamt"), col("sh_w")         # r0 = shamt + 64 * sh_w (execute.rs:287: shift = rs2 & 0x3F)
    ZFLAG = col("sh_zero")                         # right shift by 0 (the power table says so)
    G_LO, G_HI = col("fill_lo"), col("fill_hi")    # SRA: the ones shifted in for a negative value (value.rs:672-691)
    # bitwise: 5-bit pieces of the chunks of X and Y and of X & Y, looked up as triples in the AND table
    XL = [col(f"xl{i}") for i in range(4)]
    XH = [col(f"xh{i}") for i in range(4)]
    YL = [col(f"yl{i}") for i in range(4)]
    YH = [col(f"yh{i}") for i in range(4)]
    ZL = [col(f"zl{i}") for i in range(4)]
    ZH = [col(f"zh{i}") for i in range(4)]
    M_AND, M_POW = col("m_and"), col("m_pow")      # multiplicities of the AND-table row / power-table row at this trace row
    # memory (docs/PROVER_SPEC.md section 3.8): one access per row to ONE aligned 8-byte word (natural alignment, memory.rs:329,383,440).
    # The range-checked chunks ch0..ch3 hold the effective address; ch0 = offset + 8 * mw.
    OH = [col(f"off{k}") for k in range(8)]        # one-hot byte offset inside the word
    MW = col("mw")                                 # ch0 >> 3 (7 bits)
    OB = [col(f"ob{k}") for k in range(8)]         # the word's bytes before the access
    NB = [col(f"nb{k}") for k in range(8)]         # ... and after it (loads: unchanged)
    GB = [col(f"gb{k}") for k in range(5)]         # bytes of the register-side value: the loaded value / the stored register
    NL, NH = col("nib_lo"), col("nib_hi")          # gb2 = nib_lo + 16 nib_hi: limbs are 20 bits, the boundary cuts byte 2
    PTS = col("prev_ts")                           # timestamp of the previous access to the word (0 = initial value)
    TD = [col(f"td{k}") for k in range(3)]         # clk - prev_ts in three chunks: the previous access is earlier
    # boundary of the memory argument.  Image words (below the end of the code segment) get their initial value from public columns;
    # this row's cells hold the final value / timestamp of image word `row`.  RAM words (above): a prover-chosen strictly increasing
    # list, initial value zero (memory.rs:297-309), final value / timestamp committed here.
    IF = [col(f"img_fin{k}") for k in range(8)]
    IFTS = col("img_fin_ts")
    RON = col("ram_on")
    RA = [col(f"ram_a{k}") for k in range(3)]      # word index, 10 + 10 + 7 bits (addresses below 2^30)
    RE = [col(f"ram_e{k}") for k in range(3)]      # this word index - previous one - 1 (row 0: - first RAM word index)
    RF = [col(f"ram_fin{k}") for k in range(8)]
    RFTS = col("ram_fin_ts")
    M_B8, M_B4, M_B7 = col("m_b8"), col("m_b4"), col("m_b7")   # multiplicities of the 8 / 4 / 7-bit range tables
    while len(COLS) % 8:                           # a Merkle leaf absorbs whole rows in blocks of 8 columns (csrc/poseidon2.cu)
        col(f"pad{len(COLS) % 8}")                 # unconstrained, zero
WIDTH = len(COLS)
assert FULL or WIDTH == 88   # 11 sponge absorptions per Merkle leaf (rate 8)

# public columns (evaluated by the verifier): range table, ROM pc, ROM decoded word, ROM immediate
PUB_NAMES = ["p_t", "p_pc", "p_dec", "p_imm"]
P_T, P_PC, P_DEC, P_IMM = range(4)
if FULL:
    # AND table on the 1024 range rows: t = x + 32 y -> (x, y, x & y).  Power table on rows 0..127: key = s + 1024 * right (left: rows 0..63, right: rows 64..127) ->
    # (key, multiplier limbs, [right and s == 0], fill limbs): left 2^s (0 for s >= 40), right 2^(40-s) (0 for s = 0 or s > 40),
    # fill = the top min(s, 40) bits of a 40-bit word set; rows 128.. repeat row 0
    PUB_NAMES += ["p_ax", "p_ay", "p_az", "p_key", "p_mlo", "p_mhi", "p_zf", "p_glo", "p_ghi"]
    P_AX, P_AY, P_AZ, P_KEY, P_MLO, P_MHI, P_ZF, P_GLO, P_GHI = range(4, 13)
    # narrow range tables on the range rows (t below 2^k, else 0: a duplicate of row 0), and the memory image: row i = aligned word i
    # of [0, end of code): p_img_on = 1, p_img_a = i, its eight bytes; p_ram0 = first RAM word index, on row 0
    PUB_NAMES += ["p_b8", "p_b4", "p_b7", "p_img_on", "p_img_a"] + [f"p_img{k}" for k in range(8)] + ["p_ram0"]
    P_B8, P_B4, P_B7, P_IMG_ON, P_IMG_A = range(13, 18)
    P_IMG = list(range(18, 26))
    P_RAM0 = 26
PUB_WIDTH = len(PUB_NAMES)
RANGE_BITS = 10
NUM_THETA = 10 if FULL else 4
LOADS = ["s_lb", "s_lbu", "s_lh", "s_lhu", "s_lw", "s_ld"]
STORES = ["s_sb", "s_sh", "s_sw", "s_sd"]
ACCESS_WIDTH = {"s_lb": 1, "s_lbu": 1, "s_lh": 2, "s_lhu": 2, "s_lw": 4, "s_ld": 8, "s_sb": 1, "s_sh": 2, "s_sw": 4, "s_sd": 8}

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
        self.gen = 0     # bumped by flush(): reloaded columns get fresh names

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
            name = f"{self.pre}{pre}{i}" + (f"_{self.gen}" if self.gen else "")
            self.lines.append(f"  const F {name} = c.{fn}({i});  // {note}")
            self.loaded[key] = E(name, atom=True)
        return self.loaded[key]

    def PV(self, i):
        return E(f"c.PV({i})", atom=True)

    def flush(self):
        """Forget the cached column loads: later uses load again (an L1 / L2 hit) instead of keeping hundreds of values live across
        the whole constraint list -- the full profile's straight-line code would otherwise spill kilobytes per thread."""
        if FULL:
            self.loaded = {}
            self.gen += 1
            self.lines.append("  c.fence();  // the loads below stay below (register pressure of the wide profile)")

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


_DEF = None


def slice_lines(lines, roots):
    """The definitions among `lines` (`const F|X name = expr;`) that the names in `roots` depend on, in their original order."""
    import re
    defs = {}
    for k, ln in enumerate(lines):
        m = re.match(r"\s*const [FX] (\w+) = (.*?);", ln)
        if m:
            defs[m.group(1)] = (k, m.group(2))
    need, todo = set(), [r for r in roots if r in defs]
    while todo:
        n = todo.pop()
        if n in need:
            continue
        need.add(n)
        todo.extend(t for t in re.findall(r"[A-Za-z_]\w*", defs[n][1]) if t in defs and t not in need)
    return [lines[defs[n][0]] for n in sorted(need, key=lambda n: defs[n][0])]


def names_in(*exprs):
    import re
    out = []
    for e in exprs:
        out += re.findall(r"[A-Za-z_]\w*", lift(e).code)
    return out


def sum_e(xs):
    xs = list(xs)
    acc = xs[0]
    for x in xs[1:]:
        acc = acc + x
    return acc


TWO10, TWO20 = 1 << 10, 1 << 20


class LazySel:
    """Full profile: selector columns are loaded where they are used (and again after a Gen.flush()), not all 43 up front."""

    def __init__(self, g):
        self.g = g

    def __getitem__(self, n):
        return self.g.L(S[n])

    def values(self):
        return [self[n] for n in SEL_NAMES]


class LazyFam:
    """Full profile: opcode-family sums, computed on first use and again after a Gen.flush()."""

    def __init__(self, g, defs):
        self.g, self.defs = g, defs

    def __getitem__(self, k):
        key = ("fam", k)
        if key not in self.g.loaded:
            thunk, note = self.defs[k]
            self.g.loaded[key] = self.g.tmp(thunk(self), note)
        return self.g.loaded[key]


def shared(g):
    """Row-local expressions used both by the constraints and by the lookup fractions."""
    L = g.L
    s = LazySel(g) if FULL else {n: L(S[n]) for n in SEL_NAMES}
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
    rc_terms = s["s_add"] + s["s_addi"] + s["s_sub"] + L(IS_READ) + s["s_jal"] + s["s_jalr"] + s["s_sltu"] + s["s_bltu"]
    if FULL:   # signed compares (chunks = a - b mod 2^40) and the DIV family (chunks = remainder - divisor mod 2^40)
        rc_terms = rc_terms + s["s_slt"] + s["s_blt"] + s["s_divu"] + s["s_div"] + sum_e(s[n] for n in LOADS + STORES)   # memory rows: the effective address
    rc_on = g.tmp(rc_terms, "rows whose chunks are range-checked")
    opcode = g.tmp(sum_e(op * s[n] for n, op in SEL_OPCODE if op) + ECALL_OPCODE * s_ecall + L(NEG), "opcode number (opcode.rs:24-144)")
    rd_eff = g.tmp(idx(RD_H, RD_L) - 10 * (L(IS_READ) + L(IS_POS2)), "rd field of the word (READ / POSEIDON2 write r10 without an rd field)")
    dec = g.tmp(opcode + 128 * rd_eff + 2048 * idx(RS1_H, RS1_L) + 32768 * idx(RS2_H, RS2_L), "decoded word: opcode | rd << 7 | rs1 << 11 | rs2 << 15")
    imm_f = g.tmp(L(IMM_LO) - TWO20 * L(IMM_SIGN), "signed immediate as a field element (execute.rs:187)")
    fam = None
    if FULL:
        fam = LazyFam(g, dict(
            mulf=(lambda f: s["s_mul"] + 0, "MUL / MULH"),
            divf=(lambda f: s["s_divu"] + s["s_div"], "DIV family (DIV / REM act as DIVU / REMU: a provable register value is below 2^40, execute.rs:117-183)"),
            shl=(lambda f: s["s_sll"] + s["s_slli"], "left shift"),
            shr=(lambda f: s["s_srl"] + s["s_srli"] + s["s_sra"] + s["s_srai"], "right shift"),
            sraf=(lambda f: s["s_sra"] + s["s_srai"], "arithmetic right shift"),
            shi=(lambda f: s["s_slli"] + s["s_srli"] + s["s_srai"], "shift by immediate"),
            cmps=(lambda f: s["s_slt"] + s["s_blt"], "signed compare"),
            bitf=(lambda f: s["s_and"] + s["s_or"] + s["s_xor"] + s["s_andi"] + s["s_ori"] + s["s_xori"], "bitwise row"),
            shf=(lambda f: f["shl"] + f["shr"], "shift row"),
            mul_on=(lambda f: f["mulf"] + f["divf"] + f["shf"], "multiplier block active"),
            dec_on=(lambda f: f["mul_on"] + f["cmps"] + f["bitf"], "x / y chunks are range-checked"),
            r_on=(lambda f: f["divf"] + f["shf"], "r chunks are range-checked"),
            sa_on=(lambda f: f["cmps"] + f["sraf"], "sign of x is extracted"),
            ld=(lambda f: sum_e(s[n] for n in LOADS), "load row"),
            st=(lambda f: sum_e(s[n] for n in STORES), "store row"),
            mem=(lambda f: f["ld"] + f["st"], "memory row")))
    return s, s_ecall, s_pad, live, rc_on, dec, imm_f, fam


def fractions(g, sh):
    """The 8 LogUp fractions of a row as (numerator F, denominator X).  Bus 1 = 10-bit range table, bus 2 = program ROM, bus 3 = public I/O.
    Fingerprints: range `1 + theta*value`; ROM `2 + theta*pc + theta^2*dec + theta^3*imm`; I/O `3 + theta*clk + theta^2*kind + theta^3*lo + theta^4*hi`."""
    s, s_ecall, s_pad, live, rc_on, dec, imm_f, fam = sh
    L = g.L
    z = E("c.z()", atom=True, t="X")
    th = [None] + [E(f"c.th({k})", atom=True, t="X") for k in range(1, NUM_THETA + 1)]
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
    if FULL:
        def rng(n, e, note):
            out.append((n, g.tmp(z - (th[1] * e + 1), note)))
        for i in range(4):
            rng(fam["dec_on"], L(XC[i]), f"range lookup of x{i}")
        for i in range(4):
            rng(fam["dec_on"], L(YC[i]), f"range lookup of y{i}")
        for i in range(4):
            rng(fam["r_on"], L(RC[i]), f"range lookup of r{i}")
        for i in range(8):
            rng(fam["mul_on"], L(PC8[i]), f"range lookup of p{i}")
        for i in range(5):
            rng(fam["mul_on"], L(KLO[i]), f"range lookup of carry {i} (low chunk)")
            rng(fam["mul_on"], L(KHI[i]), f"range lookup of carry {i} (high chunk)")
        # bit 39: x3 = 512 * sign + rest with rest < 512  <=>  2 * rest is a 10-bit value too (x3 itself is range-checked above)
        rng(fam["sa_on"], 2 * L(XC[3]) - 1024 * L(SA), "sign of x: 2 * (x3 - 512 sign_a) in range")
        rng(fam["cmps"], 2 * L(YC[3]) - 1024 * L(SB), "sign of y: 2 * (y3 - 512 sign_b) in range")
        # shift amount: r0 = shamt + 64 w, w < 16 (shamt < 64 comes with the power-table lookup)
        rng(fam["shf"], L(SHW), "shift: w in range")
        rng(fam["shf"], 64 * L(SHW), "shift: 64 w in range (w < 16)")
        y_lo = L(YC[0]) + TWO10 * L(YC[1])
        y_hi = L(YC[2]) + TWO10 * L(YC[3])
        out.append((fam["shf"], g.tmp(z - (th[1] * (L(SHAMT) + 1024 * fam["shr"]) + th[2] * y_lo + th[3] * y_hi + th[4] * L(ZFLAG) + th[5] * L(G_LO) + th[6] * L(G_HI) + 5),
                                      "power-table lookup (shamt + 1024 right, multiplier, zero flag, fill)")))
        out.append((g.tmp(0 - L(M_POW)), g.tmp(z - (th[1] * g.Pc(P_KEY) + th[2] * g.Pc(P_MLO) + th[3] * g.Pc(P_MHI) + th[4] * g.Pc(P_ZF) + th[5] * g.Pc(P_GLO) + th[6] * g.Pc(P_GHI) + 5),
                                               "power-table row")))
        for i in range(4):
            out.append((fam["bitf"], g.tmp(z - (th[1] * L(XL[i]) + th[2] * L(YL[i]) + th[3] * L(ZL[i]) + 4), f"AND lookup, low pieces of chunk {i}")))
            out.append((fam["bitf"], g.tmp(z - (th[1] * L(XH[i]) + th[2] * L(YH[i]) + th[3] * L(ZH[i]) + 4), f"AND lookup, high pieces of chunk {i}")))
        out.append((g.tmp(0 - L(M_AND)), g.tmp(z - (th[1] * g.Pc(P_AX) + th[2] * g.Pc(P_AY) + th[3] * g.Pc(P_AZ) + 4), "AND-table row")))
        # --- memory.  Narrow range buses: 8 (byte), 9 (nibble), 10 (7 bits); memory bus 6: (word index, timestamp, the eight bytes) -- bytes as
        # separate tuple entries, so a consumer cannot re-split a packed value: every byte on the bus is a stored (range-checked) or public byte
        mem, st = fam["mem"], fam["st"]

        def nar(bus, n, e, note):
            out.append((n, g.tmp(z - (th[1] * e + bus), note)))
        for k in range(5):
            nar(8, st, L(GB[k]), f"stored byte {k} is a byte")     # loaded bytes come from memory: every byte there was checked when it was stored
        nar(9, mem, L(NL), "nibble split of byte 2 (low)")
        nar(9, mem, L(NH), "nibble split of byte 2 (high)")
        nar(10, mem, L(MW), "ch0 >> 3 has 7 bits")
        # LB / LH sign-extend to 64 bits upstream (execute.rs:477-500): a set sign bit leaves the 40-bit register model, so it must be clear
        nar(10, s["s_lb"], L(GB[0]), "lb: byte below 128")
        nar(10, s["s_lh"], L(GB[1]), "lh: high byte below 128")
        for k in range(3):
            rng(mem, L(TD[k]), f"range lookup of timestamp distance chunk {k}")
        widx = L(MW) + 128 * L(CH[1]) + (1 << 17) * L(CH[2])

        def mfp(a, ts, w):
            e = th[1] * a + sum_e(th[3 + k] * w[k] for k in range(8)) + 6
            return z - (e if ts is None else e + th[2] * ts)
        out.append((g.tmp(0 - mem), g.tmp(mfp(widx, L(PTS), [L(c) for c in OB]), "memory: consume the word as the previous access left it")))
        out.append((mem, g.tmp(mfp(widx, L(CLK) + 1, [L(c) for c in NB]), "memory: produce the word at timestamp clk + 1")))
        img_on = g.Pc(P_IMG_ON)
        out.append((img_on, g.tmp(mfp(g.Pc(P_IMG_A), None, [g.Pc(c) for c in P_IMG]), "memory: initial value of image word `row` (timestamp 0)")))
        out.append((g.tmp(0 - img_on), g.tmp(mfp(g.Pc(P_IMG_A), L(IFTS), [L(c) for c in IF]), "memory: final value of image word `row`")))
        ram_a = L(RA[0]) + TWO10 * L(RA[1]) + TWO20 * L(RA[2])
        ron = L(RON)
        out.append((ron, g.tmp(z - (th[1] * ram_a + 6), "memory: a RAM word starts as zero at timestamp 0")))
        out.append((g.tmp(0 - ron), g.tmp(mfp(ram_a, L(RFTS), [L(c) for c in RF]), "memory: final value of the RAM word")))
        rng(ron, L(RA[0]), "RAM word index chunk 0")
        rng(ron, L(RA[1]), "RAM word index chunk 1")
        nar(10, ron, L(RA[2]), "RAM word index chunk 2 (7 bits)")
        for k in range(3):
            rng(ron, L(RE[k]), f"RAM word index distance chunk {k}")
        out.append((g.tmp(0 - L(M_B8)), g.tmp(z - (th[1] * g.Pc(P_B8) + 8), "byte-table row")))
        out.append((g.tmp(0 - L(M_B4)), g.tmp(z - (th[1] * g.Pc(P_B4) + 9), "nibble-table row")))
        out.append((g.tmp(0 - L(M_B7)), g.tmp(z - (th[1] * g.Pc(P_B7) + 10), "7-bit-table row")))
    return out


# The two fractions FRAC_PHI (ROM lookup, I/O event) are added by the running-sum transition itself; the others are summed in pairs
# by the helper columns: helper k = fractions FRAC_PAIRS[k] (the last helper of an odd count holds one fraction)
FRAC_PHI = (5, 7)


def _count_fractions():
    g = Gen(pre="n_")
    return len(fractions(g, shared(g)))


def _pairs(n):
    rest = [j for j in range(n) if j not in FRAC_PHI]   # core: (0, 1), (2, 3), (4, 6)
    pairs = [tuple(rest[i:i + 2]) for i in range(0, len(rest), 2)]
    if len(pairs) % 2 == 0:   # helpers + phi = an even number of ext4 columns: the aux width stays a multiple of 8 (Merkle leaf blocks)
        pairs.append(())      # a helper that is constrained to zero
    return pairs


NUM_FRACTIONS = _count_fractions()
FRAC_PAIRS = _pairs(NUM_FRACTIONS)
# aux columns: ext4 helper sums h0.. and phi (running sum), 4 base columns each
AUX_NAMES = [f"h{k}" for k in range(len(FRAC_PAIRS))] + ["phi"]
AUX_WIDTH = 4 * len(AUX_NAMES)


def full_constraints(g, s, fam, neg, a_lo, a_hi, b_lo, b_hi, v_lo, v_hi, rc_lo, rc_hi, k0, k1):
    """Full profile: multiplier block, signed compares, shifts, bitwise operations (docs/PROVER_SPEC.md section 3.7)."""
    L = g.L
    x = [L(c) for c in XC]
    y = [L(c) for c in YC]
    r = [L(c) for c in RC]
    p = [L(c) for c in PC8]
    kk = [g.tmp(L(KLO[i]) + TWO10 * L(KHI[i]), f"carry {i}") for i in range(5)]
    mulf, divf, shf, cmps, bitf, mul_on = (fam[k] for k in ("mulf", "divf", "shf", "cmps", "bitf", "mul_on"))
    x_lo, x_hi = g.tmp(x[0] + TWO10 * x[1], "x.lo"), g.tmp(x[2] + TWO10 * x[3], "x.hi")
    y_lo, y_hi = g.tmp(y[0] + TWO10 * y[1], "y.lo"), g.tmp(y[2] + TWO10 * y[3], "y.hi")
    r_lo, r_hi = g.tmp(r[0] + TWO10 * r[1], "r.lo"), g.tmp(r[2] + TWO10 * r[3], "r.hi")
    pl_lo, pl_hi = g.tmp(p[0] + TWO10 * p[1], "low product word, lo limb"), g.tmp(p[2] + TWO10 * p[3], "low product word, hi limb")
    ph_lo, ph_hi = g.tmp(p[4] + TWO10 * p[5], "high product word, lo limb"), g.tmp(p[6] + TWO10 * p[7], "high product word, hi limb")
    # --- operand binding
    x_is_a = g.tmp(mulf + shf + cmps + bitf, "x = rs1")
    g.emit(x_is_a * (a_lo - x_lo), "x.lo = a.lo")
    g.emit(x_is_a * (a_hi - x_hi), "x.hi = a.hi")
    y_is_b = g.tmp(mulf + divf + cmps + bitf, "y = rs2 (or the immediate)")
    g.emit(y_is_b * (b_lo - y_lo), "y.lo = b.lo")
    g.emit(y_is_b * (b_hi - y_hi), "y.hi = b.hi")
    g.emit(shf * (b_lo - r_lo), "shift: r.lo = b.lo")
    g.emit(shf * (b_hi - r_hi), "shift: r.hi = b.hi")
    # --- X * Y + R = P: the five low base-2^10 columns with carries (the identity mod 2^50) and the identity in the field; both
    # sides are below 2^80 < p * 2^50, so the two congruences give equality over the integers.  Every column equation stays below p:
    # four products < 2^22, carry < 2^20, 1024 * carry < 2^30
    for k in range(5):
        sk = sum_e(x[i] * y[k - i] for i in range(4) if 0 <= k - i < 4)
        prev = kk[k - 1] if k else 0
        radd = divf * r[k] if k < 4 else 0
        g.emit(mul_on * (sk + prev - p[k] - TWO10 * kk[k]) + radd, f"multiplier column {k}")
    xf = g.tmp(x_lo + TWO20 * x_hi, "x as a field element")
    yf = g.tmp(y_lo + TWO20 * y_hi, "y as a field element")
    pf = g.tmp(sum_e(pow(2, 10 * k, P) * p[k] for k in range(8)), "the 80-bit product as a field element")
    g.emit(mul_on * (xf * yf - pf) + divf * (r_lo + TWO20 * r_hi), "multiplier identity mod p")
    # --- MUL / MULH (execute.rs:80-106): low / high 40 bits of the product
    g.emit(mulf * (v_lo - pl_lo - neg * (ph_lo - pl_lo)), "mul / mulh result lo")
    g.emit(mulf * (v_hi - pl_hi - neg * (ph_hi - pl_hi)), "mul / mulh result hi")
    # --- DIVU / REMU / DIV / REM (execute.rs:117-183): a = q * b + rem with rem < b (so b != 0: a zero divisor is a VM error, no row)
    g.emit(divf * (a_lo - pl_lo), "div: q * b + rem = a (lo)")
    g.emit(divf * (a_hi - pl_hi), "div: q * b + rem = a (hi)")
    g.emit(divf * ph_lo, "div: no overflow (lo)")
    g.emit(divf * ph_hi, "div: no overflow (hi)")
    g.emit(divf * (v_lo - x_lo - neg * (r_lo - x_lo)), "quotient / remainder result lo")
    g.emit(divf * (v_hi - x_hi - neg * (r_hi - x_hi)), "quotient / remainder result hi")
    g.emit(divf * (r_lo - b_lo - rc_lo + TWO20 * k0), "rem - b, lo limb (the chunks hold rem - b mod 2^40)")
    g.emit(divf * (r_hi - b_hi - k0 - rc_hi + TWO20 * k1), "rem - b, hi limb")
    g.emit(divf * (k1 - 1), "rem < b")
    # --- signed compares at bit 39 (execute.rs:361-391, 594-609): (a <s b) = (a <u b) xor sign_a xor sign_b
    g.flush()
    cmps, shf, shl, shr, sraf, shi = (fam[k] for k in ("cmps", "shf", "shl", "shr", "sraf", "shi"))
    r = [L(c) for c in RC]
    p = [L(c) for c in PC8]
    pl_lo, pl_hi = g.tmp(p[0] + TWO10 * p[1], "low product word, lo limb"), g.tmp(p[2] + TWO10 * p[3], "low product word, hi limb")
    ph_lo, ph_hi = g.tmp(p[4] + TWO10 * p[5], "high product word, lo limb"), g.tmp(p[6] + TWO10 * p[7], "high product word, hi limb")
    sa, sb, sx, lts = L(SA), L(SB), L(SXOR), L(LTS)
    g.emit(sx - (sa + sb - 2 * sa * sb), "sign_xor = sign_a xor sign_b")
    g.emit(cmps * (lts - (k1 + sx - 2 * k1 * sx)), "lt_signed = borrow xor sign_xor")
    g.emit(s["s_slt"] * (v_lo - (lts + neg - 2 * lts * neg)), "slt / sge result")
    g.emit(s["s_slt"] * v_hi, "slt / sge result is 0 / 1")
    # --- shifts (execute.rs:282-358, value.rs:658-691): x * 2^s -> low word, x * 2^(40-s) -> high word = x >> s; the multiplier,
    # the "shift by zero" flag and the arithmetic fill come from the power table, keyed by shamt (+ 64 for right shifts)
    g.emit(shf * (r[0] - L(SHAMT) - 64 * L(SHW)), "shift amount = low 6 bits of b")
    g.emit(shi * L(SHW), "immediate shifts: shamt < 64")
    g.emit(shl * (v_lo - pl_lo), "left shift result lo")
    g.emit(shl * (v_hi - pl_hi), "left shift result hi")
    zf = L(ZFLAG)
    g.emit(shr * (v_lo - ph_lo - zf * a_lo) - sraf * sa * L(G_LO), "right shift result lo (+ sign fill)")
    g.emit(shr * (v_hi - ph_hi - zf * a_hi) - sraf * sa * L(G_HI), "right shift result hi (+ sign fill)")
    # --- bitwise (execute.rs:200-279): chunks split into 5-bit pieces, z = x & y piecewise from the AND table; or = x + y - z, xor = x + y - 2 z
    g.flush()
    bitf = fam["bitf"]
    x = [L(c) for c in XC]
    y = [L(c) for c in YC]
    zc = []
    for i in range(4):
        g.emit(bitf * (x[i] - L(XL[i]) - 32 * L(XH[i])), f"x{i} = low + 32 high piece")
        g.emit(bitf * (y[i] - L(YL[i]) - 32 * L(YH[i])), f"y{i} = low + 32 high piece")
        zc.append(L(ZL[i]) + 32 * L(ZH[i]))
    z_lo, z_hi = g.tmp(zc[0] + TWO10 * zc[1], "(x & y).lo"), g.tmp(zc[2] + TWO10 * zc[3], "(x & y).hi")
    andf, orf, xorf = s["s_and"] + s["s_andi"], s["s_or"] + s["s_ori"], s["s_xor"] + s["s_xori"]
    g.emit(bitf * (v_lo - a_lo - b_lo) + andf * (a_lo + b_lo - z_lo) + orf * z_lo + xorf * 2 * z_lo, "bitwise result lo")
    g.emit(bitf * (v_hi - a_hi - b_hi) + andf * (a_hi + b_hi - z_hi) + orf * z_hi + xorf * 2 * z_hi, "bitwise result hi")


def memory_constraints(g, s, fam, a_lo, a_hi, b_lo, b_hi, v_lo, v_hi, rc_lo, rc_hi, k0, k1, trans, first):
    """Full profile: loads and stores (execute.rs:477-575, memory.rs:297-489) as one access to an aligned 8-byte word, checked offline
    (docs/PROVER_SPEC.md section 3.8): every access consumes (word, previous timestamp, bytes) and produces (word, clk + 1, bytes)."""
    L, N = g.L, g.N
    mem, ld, st = fam["mem"], fam["ld"], fam["st"]
    oh = [L(c) for c in OH]
    ob = [L(c) for c in OB]
    nb = [L(c) for c in NB]
    gb = [L(c) for c in GB]
    ch = [L(c) for c in CH]
    for k in range(8):
        g.emit(oh[k] * (oh[k] - 1), f"bool off{k}")
    g.emit(sum_e(oh) - mem, "exactly one byte offset on a memory row, none elsewhere")
    # effective address = rs1 + imm, no wrap in 40 bits (execute.rs:478 wraps in 64: such an address is not provable), below 2^30
    imm_hi = (TWO20 - 1) * L(IMM_SIGN)
    g.emit(mem * (a_lo + L(IMM_LO) - rc_lo - TWO20 * k0), "address lo limb")
    g.emit(mem * (a_hi + imm_hi + k0 - rc_hi - TWO20 * k1), "address hi limb")
    g.emit(mem * (k1 - L(IMM_SIGN)), "address: a negative offset borrows exactly once, a positive one never carries out")
    g.emit(mem * ch[3], "address below 2^30")
    g.emit(mem * (ch[0] - sum_e(k * oh[k] for k in range(1, 8)) - 8 * L(MW)), "ch0 = byte offset + 8 * mw")
    # natural alignment (memory.rs:329,383,440: MisalignedAccess is a VM error, no row)
    for w in (2, 4, 8):
        sel = sum_e(s[n] for n, ww in ACCESS_WIDTH.items() if ww == w)
        g.emit(sel * sum_e(oh[k] for k in range(8) if k % w), f"{w}-byte accesses are aligned")
    # the previous access is earlier: clk + 1 - prev_ts - 1 >= 0
    g.emit(mem * (L(CLK) - L(PTS) - L(TD[0]) - TWO10 * L(TD[1]) - TWO20 * L(TD[2])), "clk - prev_ts is a 30-bit value")
    # register-side value g: bytes gb0..gb4 (byte 2 split into nibbles at the limb boundary)
    g.emit(mem * (gb[2] - L(NL) - 16 * L(NH)), "byte 2 = nibbles")
    g_lo = g.tmp(gb[0] + 256 * gb[1] + 65536 * L(NL), "register-side value, lo limb")
    g_hi = g.tmp(L(NH) + 16 * gb[3] + 4096 * gb[4], "register-side value, hi limb")
    g.emit(ld * (v_lo - g_lo) + st * (b_lo - g_lo), "loaded value / stored register, lo limb")
    g.emit(ld * (v_hi - g_hi) + st * (b_hi - g_hi), "loaded value / stored register, hi limb")
    # loads (zero-extended; LB / LH with a clear sign bit; LD needs bytes 5..7 zero to stay inside 40 bits): gb[j] = ob[offset + j] for j < width
    for j in range(5):
        terms = []
        for n in LOADS:
            w = ACCESS_WIDTH[n]
            if j < w:
                terms.append(s[n] * sum_e(oh[k] * ob[k + j] for k in range(0, 8, w)))
        g.emit(ld * gb[j] - sum_e(terms), f"load: byte {j} of the value")
    for k in (5, 6, 7):
        g.emit(s["s_ld"] * ob[k], f"ld: byte {k} is zero (the value fits 40 bits)")
    # new bytes: stores replace `width` bytes at the offset with the low bytes of rs2 (bytes 5..7 of an SD are zero), loads change nothing
    for k in range(8):
        terms = []
        for n in STORES:
            w = ACCESS_WIDTH[n]
            base = k - k % w
            j = k - base
            src = gb[j] if j < 5 else 0
            terms.append(s[n] * oh[base] * (src - ob[k]))
        g.emit(mem * (nb[k] - ob[k]) - sum_e(terms), f"new byte {k}")
    # --- boundary: RAM list = strictly increasing word indices from p_ram0 on, packed at the first rows
    ron = L(RON)
    g.emit(ron * (ron - 1), "bool ram_on")
    g.emit(trans * (N(RON) * (1 - ron)), "the RAM list is packed at the first rows")
    ram_a = g.tmp(L(RA[0]) + TWO10 * L(RA[1]) + TWO20 * L(RA[2]), "RAM word index")
    ram_an = N(RA[0]) + TWO10 * N(RA[1]) + TWO20 * N(RA[2])
    dist = L(RE[0]) + TWO10 * L(RE[1]) + TWO20 * L(RE[2])
    dist_n = N(RE[0]) + TWO10 * N(RE[1]) + TWO20 * N(RE[2])
    g.emit(trans * (N(RON) * (ram_an - ram_a - 1 - dist_n)), "RAM word indices increase strictly")
    g.emit(first * (ron * (ram_a - g.Pc(P_RAM0) - dist)), "the first RAM word lies above the program image")


def logup_full(g, sh, first, last, trans, sio):
    """LogUp constraints of the full profile.  Same constraints as the core tail of build(), but every helper is evaluated inside its
    own block from freshly loaded cells (the definitions it needs are sliced out of a scratch generator), and the helpers are summed
    as they go: 81 ext4 denominators never have to be live at once."""
    gs = Gen(pre="q")
    fr = fractions(gs, sh)
    NH = len(FRAC_PAIRS)

    def xaux(k, nxt=False):
        ld = gs.AN if nxt else gs.A
        cs = [ld(4 * k + j) for j in range(4)]
        return gs.tmp(E(f"c.x4({cs[0].code}, {cs[1].code}, {cs[2].code}, {cs[3].code})", atom=True, t="X"), ("next " if nxt else "") + AUX_NAMES[k])
    h = [xaux(k) for k in range(NH)]
    phi, phi_n = xaux(NH), xaux(NH, True)
    g.lines.append("  X hsum = c.xf(c.K(0u));  // sum of the helpers, accumulated block by block")
    for k, pair in enumerate(FRAC_PAIRS):
        if len(pair) == 0:
            expr, note = h[k], f"helper {k} pads the aux width: zero"
        elif len(pair) == 1:
            (ni, di), = (fr[pair[0]],)
            expr, note = h[k] * di - ni, f"helper {k} = fraction {pair[0]}"
        else:
            i, j = pair
            (ni, di), (nj, dj) = fr[i], fr[j]
            expr, note = h[k] * di * dj - di * nj - dj * ni, f"helper {k} = fraction {i} + fraction {j}"
        g.lines.append("  {")
        g.lines.extend(slice_lines(gs.lines, names_in(expr)))
        g.emit(expr, note)
        g.lines.append(f"  hsum = hsum + {h[k].code};")
        g.lines.append("  }")
        if k % 2:
            g.lines.append("  c.fence();")
    (n5, d5), (n7, d7) = fr[FRAC_PHI[0]], fr[FRAC_PHI[1]]
    hs = E("hsum", atom=True, t="X")
    tail = [((phi_n - phi - hs) * d5 * d7 - d7 * n5 - d5 * n7) * trans, phi * first, (sio - phi - hs) * last]
    g.lines.append("  {")
    g.lines.extend(slice_lines(gs.lines, names_in(*tail)))
    g.emit(tail[0], "running sum transition (adds the ROM lookup and the I/O event itself)")
    g.emit(tail[1], "running sum starts at 0")
    g.emit(tail[2], "range and ROM lookups cancel; what remains is the public I/O transcript's sum")
    g.lines.append("  }")


def build():
    g = Gen()
    L, N = g.L, g.N
    first, last, trans = E("c.is_first", True), E("c.is_last", True), E("c.is_trans", True)
    sh = shared(g)
    s, s_ecall, s_pad, live, rc_on, dec, imm_f, fam = sh
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
    paired = s["s_beq"] + s["s_sltu"] + s["s_seq"] + s["s_bltu"] + s["s_cmov"]
    if FULL:
        paired = paired + s["s_mul"] + s["s_divu"] + s["s_div"] + s["s_slt"] + s["s_blt"]
        for b in (SA, SB):
            g.emit(L(b) * (L(b) - 1), f"bool {COLS[b]}")
    g.emit(neg * (1 - paired), "polarity only on the paired families")

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
    immb = s["s_addi"]
    if FULL:   # ANDI / ORI / XORI use the sign-extended immediate (execute.rs:241-279), the immediate shifts their shamt (:316-358)
        immb = g.tmp(s["s_addi"] + s["s_andi"] + s["s_ori"] + s["s_xori"] + fam["shi"], "b is the immediate")
    g.emit(b_lo - rs2_lo - immb * L(IMM_LO), "b.lo = reg[rs2].lo + [immediate form] * imm.lo")
    imm_hi = g.tmp((TWO20 - 1) * L(IMM_SIGN), "imm.hi = sign-extension limb")
    g.emit(b_hi - rs2_hi - immb * imm_hi, "b.hi = reg[rs2].hi + [immediate form] * imm.hi")
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
    cmpu = s["s_sltu"] + s["s_bltu"]
    if FULL:
        cmpu = cmpu + fam["cmps"]
    cmpu = g.tmp(cmpu, "compare row: the chunks hold a - b mod 2^40")
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
    if FULL:
        g.flush()
        full_constraints(g, s, fam, neg, a_lo, a_hi, b_lo, b_hi, v_lo, v_hi, rc_lo, rc_hi, k0, k1)
        g.flush()
        memory_constraints(g, s, fam, a_lo, a_hi, b_lo, b_hi, v_lo, v_hi, rc_lo, rc_hi, k0, k1, trans, first)
    g.flush()
    # --- syscalls (syscall.rs:94-149)
    g.emit((L(IS_READ) + L(IS_POS2)) * (rd_h[2] * rd_l[2] - 1), "read / poseidon2 write r10 (syscall.rs:104-109,140-149)")
    g.emit(L(IS_WRITE) * (v_lo - L(REG_LO[11])), "write: v = the written word r11 (lo), sent to the I/O bus (syscall.rs:110-119)")
    g.emit(L(IS_WRITE) * (v_hi - L(REG_HI[11])), "write: v = r11 (hi)")
    g.emit(L(IS_POS2) * v_lo, "poseidon2 returns 0 (lo)")
    g.emit(L(IS_POS2) * v_hi, "poseidon2 returns 0 (hi)")
    # --- register write-back, pre-state rows: next.r[i] = (rd == i && w) ? v : r[i]
    w = addlike + s["s_sub"] + s["s_jal"] + s["s_jalr"] + s["s_sltu"] + s["s_seq"] + L(IS_READ) + L(IS_POS2) + cm * mv
    if FULL:
        w = w + fam["mul_on"] + fam["bitf"] + s["s_slt"] + fam["ld"]
    w = g.tmp(w, "write enable")
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
    br = s["s_beq"] + s["s_bltu"]
    if FULL:
        lts = L(LTS)
        g.emit(s["s_blt"] * (taken - (lts + neg - 2 * lts * neg)), "blt / bge taken (execute.rs:594-609)")
        br = br + s["s_blt"]
    br = g.tmp(br, "branch row")
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
    # --- LogUp: helper k = n_i/d_i + n_j/d_j; phi' = phi + sum of the helpers + n_5/d_5 + n_7/d_7; phi_first = 0; the last row closes the sum
    sio = E("c.sio()", atom=True, t="X")
    if FULL:
        logup_full(g, sh, first, last, trans, sio)
        return g
    fr = fractions(g, sh)

    def xaux(k, nxt=False):
        ld = g.AN if nxt else g.A
        cs = [ld(4 * k + j) for j in range(4)]
        return g.tmp(E(f"c.x4({cs[0].code}, {cs[1].code}, {cs[2].code}, {cs[3].code})", atom=True, t="X"), ("next " if nxt else "") + AUX_NAMES[k])
    NH = len(FRAC_PAIRS)
    h = [xaux(k) for k in range(NH)]
    phi, phi_n = xaux(NH), xaux(NH, True)
    for k, pair in enumerate(FRAC_PAIRS):
        i, j = pair
        (ni, di), (nj, dj) = fr[i], fr[j]
        g.emit(h[k] * di * dj - di * nj - dj * ni, f"helper {k} = fraction {i} + fraction {j}")
    (n5, d5), (n7, d7) = fr[FRAC_PHI[0]], fr[FRAC_PHI[1]]
    hs = g.tmp(sum_e(h), "h0 + h1 + h2")
    g.emit(((phi_n - phi - hs) * d5 * d7 - d7 * n5 - d5 * n7) * trans, "running sum transition (adds the ROM lookup and the I/O event itself)")
    g.emit(phi * first, "running sum starts at 0")
    # last row = padding row: its ROM-lookup and I/O numerators are 0 (live = 0), only the table-side helpers count
    g.emit((sio - phi - hs) * last, "range and ROM lookups cancel; what remains is the public I/O transcript's sum")
    return g


def build_fractions():
    g = Gen(pre="f_")
    sh = shared(g)
    if not FULL:
        fr = fractions(g, sh)
        for j, (n, d) in enumerate(fr):
            g.lines.append(f"  c.frac({j}, {lift(n).code}, {d.code});")
        return g
    gs = Gen(pre="fq")   # full profile: one block per fraction (see logup_full); the context consumes fraction j before j + 1 is built
    fr = fractions(gs, sh)
    for j, (n, d) in enumerate(fr):
        g.lines.append("  {")
        g.lines.extend(slice_lines(gs.lines, names_in(n, d)))
        g.lines.append(f"  c.frac({j}, {lift(n).code}, {d.code});")
        g.lines.append("  }")
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
    hdr.append(f"#define ZKIR_AIR_NUM_HELPERS {len(FRAC_PAIRS)}")
    hdr.append(f"#define ZKIR_AIR_NUM_THETA {NUM_THETA}")
    hdr.append("#define ZKIR_AIR_MAX_DEGREE 3")
    hdr.append("// fraction j is summed by helper ZKIR_AIR_FRAC_HELPER[j] (ZKIR_AIR_NUM_HELPERS = added by the running-sum transition itself)")
    helper_of = [len(FRAC_PAIRS)] * NUM_FRACTIONS   # number of helpers = the running sum itself
    for k, pair in enumerate(FRAC_PAIRS):
        for i in pair:
            helper_of[i] = k
    hdr.append("#define ZKIR_AIR_FRAC_HELPER_INIT {" + ", ".join(str(x) for x in helper_of) + "}")
    hdr.append("#ifndef ZKIR_HD\n#ifdef __CUDACC__\n#define ZKIR_HD __host__ __device__ __forceinline__\n#else\n#define ZKIR_HD inline\n#endif\n#endif")
    hdr.append("// Context contract: C::F (base) and C::X (ext4) with + - *, X * F scaling; c.L(i)/c.N(i) main local/next row, c.A(i)/c.AN(i) aux,")
    hdr.append("// c.P(i) public column, c.PV(i) public value, c.K(u32 canonical constant), c.z()/c.th(k) lookup challenges z, theta^k,")
    hdr.append("// c.xf(F) -> X, c.x4(F,F,F,F) -> X, c.sio() -> X (sum of the public I/O transcript's fractions), c.is_first / c.is_last / c.is_trans selectors,")
    hdr.append("// c.emit(index, F), c.emit_x(index, X); full profile: c.fence() = a scheduling barrier for loads (a no-op off the GPU).")
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
    sfx = "_full" if FULL else ""
    for rel in (f"zkir_b200/csrc/air_generated{sfx}.h", f"oracle/air_generated{sfx}.h"):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(text)
    # column map: C header for the packer + python module for tests
    ch = ["// GENERATED by tools/gen_air.py -- column indices of the AIR v2.", "#pragma once"]
    for i, n in enumerate(COLS):
        ch.append(f"#define ZKIR_COL_{n.upper()} {i}")
    ch.append(f"#define ZKIR_COL_COUNT {WIDTH}")
    for i, n in enumerate(PUB_NAMES):
        ch.append(f"#define ZKIR_PUB_{n.upper()} {i}")
    with open(os.path.join(ROOT, f"zkir_b200/csrc/air_columns{sfx}.h"), "w") as f:
        f.write("\n".join(ch) + "\n")
    with open(os.path.join(ROOT, f"zkir_b200/air_layout{sfx}.py"), "w") as f:
        f.write('"""GENERATED by tools/gen_air.py -- column map of the AIR v2."""\n')
        f.write(f"WIDTH = {WIDTH}\nAUX_WIDTH = {AUX_WIDTH}\nPUB_WIDTH = {PUB_WIDTH}\nNUM_CONSTRAINTS = {g.idx}\nNUM_PUBLIC = {NUM_PUBLIC}\n")
        f.write(f"MIN_LOG_N = {RANGE_BITS}\n")
        f.write(f"PUBLIC_NAMES = {PV_NAMES!r}\n")
        f.write("COLUMNS = " + repr(COLS) + "\n")
        f.write("INDEX = {n: i for i, n in enumerate(COLUMNS)}\n")
    print(f"AIR {'full' if FULL else 'core'}: width={WIDTH} aux={AUX_WIDTH} pub={PUB_WIDTH} constraints={g.idx} fractions={NUM_FRACTIONS}")
    named = {n: i for i, n in enumerate(COLS)}
    return dict(width=WIDTH, aux=AUX_WIDTH, pub=PUB_WIDTH, constraints=g.idx, fractions=NUM_FRACTIONS, theta=NUM_THETA,
                cols={k: named[k] for k in ("img_fin0", "ram_fin_ts", "m_b7") if k in named})


def write_profiles(core, full):
    """Numbers of both profiles for the code that serves either (prover.cu, proof_layout.h): which profile a proof uses is its width."""
    out = ["// GENERATED by tools/gen_air.py -- the two AIR profiles (docs/PROVER_SPEC.md section 3.7).", "#pragma once"]
    for name, d in (("CORE", core), ("FULL", full)):
        for k in ("width", "aux", "pub", "constraints", "theta"):
            out.append(f"#define ZKIR_PROFILE_{name}_{k.upper()} {d[k]}")
    # the boundary cells of the memory argument are one contiguous column range (img_fin0 .. ram_fin_ts): the prover uploads them as a block
    out.append(f"#define ZKIR_PROFILE_FULL_COL_BOUNDARY0 {full['cols']['img_fin0']}")
    out.append(f"#define ZKIR_PROFILE_FULL_BOUNDARY_COLS {full['cols']['ram_fin_ts'] - full['cols']['img_fin0'] + 1}")
    out.append(f"#define ZKIR_PROFILE_FULL_COL_M_B7 {full['cols']['m_b7']}")
    out.append(f"#define ZKIR_PROFILE_MAX_CONSTRAINTS {max(core['constraints'], full['constraints'])}")
    out.append(f"#define ZKIR_PROFILE_MAX_AUX {max(core['aux'], full['aux'])}")
    out.append(f"#define ZKIR_PROFILE_MAX_PUB {max(core['pub'], full['pub'])}")
    text = "\n".join(out) + "\n"
    for rel in ("zkir_b200/csrc/air_profiles_generated.h", "oracle/air_profiles_generated.h"):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(text)


if __name__ == "__main__":
    import json
    import subprocess
    import sys
    if "--one" in sys.argv:      # one profile (ZKIR_AIR_PROFILE), summary as the last stdout line
        print(json.dumps(main()))
    else:
        res = {}
        for prof in ("core", "full"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(os.environ, ZKIR_AIR_PROFILE=prof), capture_output=True, text=True, check=True)
            lines = r.stdout.strip().splitlines()
            print(lines[0])
            res[prof] = json.loads(lines[-1])
        write_profiles(res["core"], res["full"])
