"""ZKIR v3.4 instruction helpers on the host side: opcode table, encode/decode (through the C ABI), a minimal text
assembler for the v3.4 syntax the reference's tests use, and the Program container with its 32-byte header.

Mirrors: zkir-spec/src/opcode.rs:24-144, zkir-assembler/src/encoder.rs:18-151, zkir-disassembler/src/decoder.rs:20-192,
zkir-assembler/src/assembler.rs:236-322 + parser.rs:16-48 (mnemonics, ABI aliases of the *assembler* table),
zkir-spec/src/program.rs:37-40,62-95,170-213,300-346 (header layout, to_bytes/from_bytes).
"""
import struct
from . import _ffi

OPCODES = {
    "add": 0x00, "sub": 0x01, "mul": 0x02, "mulh": 0x03, "divu": 0x04, "remu": 0x05, "div": 0x06, "rem": 0x07, "addi": 0x08,
    "and": 0x10, "or": 0x11, "xor": 0x12, "andi": 0x13, "ori": 0x14, "xori": 0x15,
    "sll": 0x18, "srl": 0x19, "sra": 0x1A, "slli": 0x1B, "srli": 0x1C, "srai": 0x1D,
    "sltu": 0x20, "sgeu": 0x21, "slt": 0x22, "sge": 0x23, "seq": 0x24, "sne": 0x25, "cmov": 0x26, "cmovz": 0x27, "cmovnz": 0x28,
    "lb": 0x30, "lbu": 0x31, "lh": 0x32, "lhu": 0x33, "lw": 0x34, "ld": 0x35,
    "sb": 0x38, "sh": 0x39, "sw": 0x3A, "sd": 0x3B,
    "beq": 0x40, "bne": 0x41, "blt": 0x42, "bge": 0x43, "bltu": 0x44, "bgeu": 0x45,
    "jal": 0x48, "jalr": 0x49, "ecall": 0x50, "ebreak": 0x51,
}
MNEMONIC = {v: k for k, v in OPCODES.items()}
R_TYPE = {n for n, o in OPCODES.items() if o <= 0x07 or 0x10 <= o <= 0x12 or 0x18 <= o <= 0x1A or 0x20 <= o <= 0x28}
I_TYPE = {"addi", "andi", "ori", "xori", "slli", "srli", "srai", "jalr"}
LOADS = {"lb", "lbu", "lh", "lhu", "lw", "ld"}
STORES = {"sb", "sh", "sw", "sd"}
BRANCHES = {"beq", "bne", "blt", "bge", "bltu", "bgeu"}
# the assembler's alias table (zkir-assembler/src/parser.rs:16-48); the syscall ABI itself is by number (R10..R13)
ALIASES = {"zero": 0, "ra": 1, "sp": 2, "gp": 3, "tp": 4, "fp": 5, "s0": 6, "s1": 7, "t0": 8, "t1": 9, "t2": 10,
           "a0": 11, "a1": 12, "a2": 13, "a3": 14, "a4": 15}

MAGIC = 0x52494B5A  # "ZKIR" little endian
VERSION = 0x00030004
CODE_BASE = 0x1000


def encode(mnemonic, a=0, b=0, c=0, imm=0):
    """encode('add', rd, rs1, rs2) | encode('addi', rd, rs1, imm=..) | encode('sw', rs1, rs2, imm=..) |
    encode('bne', rs1, rs2, imm=offset) | encode('jal', rd, imm=offset) | encode('ecall')"""
    return _ffi.lib().zkir_encode(OPCODES[mnemonic], a, b, c, imm)


def decode(word):
    """-> (mnemonic, a, b, c, imm) with the field meaning of encode(); raises ValueError on an unknown opcode."""
    out = (_ffi.C.c_uint32 * 5)()
    if _ffi.lib().zkir_decode(word & 0xFFFFFFFF, out) != 0:
        raise ValueError(f"UnknownOpcode({word & 0x7F:#x})")
    imm = out[4] - (1 << 32) if out[4] & 0x80000000 else out[4]
    return MNEMONIC[out[0]], out[1], out[2], out[3], imm


def _reg(tok):
    t = tok.strip().lower()
    if t in ALIASES:
        return ALIASES[t]
    if t.startswith("r") and t[1:].isdigit() and 0 <= int(t[1:]) < 16:
        return int(t[1:])
    raise ValueError(f"invalid register: {tok!r}")


def _imm(tok):
    return int(tok.strip(), 0)


def assemble_line(line):
    toks = line.replace(",", " ").replace("(", " ").replace(")", " ").split()
    m = toks[0].lower()
    ops = toks[1:]
    if m not in OPCODES:
        raise ValueError(f"unknown mnemonic: {m!r}")
    if m in ("ecall", "ebreak"):
        return encode(m)
    if m in R_TYPE:
        return encode(m, _reg(ops[0]), _reg(ops[1]), _reg(ops[2]))
    if m in I_TYPE:
        return encode(m, _reg(ops[0]), _reg(ops[1]), imm=_imm(ops[2]))
    if m in LOADS:      # lw rd, imm(rs1)
        return encode(m, _reg(ops[0]), _reg(ops[2]), imm=_imm(ops[1]))
    if m in STORES:     # sw rs2, imm(rs1): the value register comes first in the text, the base first in the word
        return encode(m, _reg(ops[2]), _reg(ops[0]), imm=_imm(ops[1]))
    if m in BRANCHES:
        return encode(m, _reg(ops[0]), _reg(ops[1]), imm=_imm(ops[2]))
    if m == "jal":
        return encode(m, _reg(ops[0]), imm=_imm(ops[1]))
    raise ValueError(m)


class Program:
    """zkir-spec/src/program.rs:241-250: header + code words + data bytes."""

    def __init__(self, code=(), data=b"", entry_point=CODE_BASE, limb_bits=20, data_limbs=2, addr_limbs=2):
        self.code = [int(w) & 0xFFFFFFFF for w in code]
        self.data = bytes(data)
        self.entry_point = entry_point
        self.limb_bits, self.data_limbs, self.addr_limbs = limb_bits, data_limbs, addr_limbs
        self.flags = self.bss_size = self.stack_size = 0

    def to_bytes(self):
        hdr = struct.pack("<IIBBBBIIIII", MAGIC, VERSION, self.limb_bits, self.data_limbs, self.addr_limbs, self.flags,
                          self.entry_point, len(self.code) * 4, len(self.data), self.bss_size, self.stack_size)
        assert len(hdr) == 32
        return hdr + struct.pack(f"<{len(self.code)}I", *self.code) + self.data

    @classmethod
    def from_bytes(cls, b):
        if len(b) < 32:
            raise ValueError("InvalidHeaderSize")
        magic, version, lb, dl, al, flags, entry, csz, dsz, bss, stk = struct.unpack("<IIBBBBIIIII", b[:32])
        if magic != MAGIC:
            raise ValueError(f"InvalidMagic({magic:#x})")
        if version != VERSION:
            raise ValueError(f"UnsupportedVersion({version:#x})")
        if not (16 <= lb <= 30 and lb % 2 == 0 and 1 <= dl <= 4 and 1 <= al <= 2):
            raise ValueError("InvalidConfig")
        if csz % 4 or len(b) != 32 + csz + dsz:
            raise ValueError("SizeMismatch")
        p = cls(struct.unpack(f"<{csz // 4}I", b[32:32 + csz]), b[32 + csz:], entry, lb, dl, al)
        p.flags, p.bss_size, p.stack_size = flags, bss, stk
        return p


def assemble(source):
    """Assemble v3.4 text (one instruction per line, `#` comments, numeric branch offsets) into a Program."""
    code = []
    for raw in source.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line or line.endswith(":") or line.startswith("."):
            continue
        code.append(assemble_line(line))
    return Program(code)
