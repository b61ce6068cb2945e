// Row converter of the AIR v2: one interpreter row (pc, instruction word, PRE-state registers) -> the per-row main columns
// (the LogUp multiplicity columns are histograms over all rows and are filled by the callers).  Core profile: 86 columns; the full
// profile (ZKIR_PROFILE_FULL, host packer only) adds the multiplier block, the sign / shift cells and the 5-bit pieces of the
// bitwise operations (tools/gen_air.py, docs/PROVER_SPEC.md section 3.7).
//
// The reference names this step but does not contain it (zkir-spec/src/trace.rs:41 "converter", zkir-runtime/src/vm.rs:243-244).
// ONE definition shared by the host packer (host/pack.cc) and the device converter (trace_expand.cu): `__host__ __device__`,
// no dynamic indexing of the register array, output through a writer functor.  Column meaning: tools/gen_air.py.
// Semantics restated from zkir-runtime/src/execute.rs (ADD :43-63, SUB :65-78, ADDI :185-197, SLTU/SGEU/SEQ/SNE :330-420,
// CMOV* :422-470, BEQ/BNE/BLTU/BGEU :578-637, JAL/JALR :639-659, ECALL/EBREAK :661-673) and syscall.rs:94-149; full profile:
// MUL/MULH :80-106, DIVU/REMU/DIV/REM :117-183, AND..XORI :200-279, SLL..SRAI :282-358 (zkir-spec/src/value.rs:658-691),
// SLT/SGE :361-391, BLT/BGE :594-609.
#pragma once
#include <stdint.h>
#include "bb.cuh"
#include "air_profile.h"

namespace zkir {

enum PackErr : u32 {
  PACK_OK = 0, PACK_ERR_PC = 1, PACK_ERR_REG40 = 2, PACK_ERR_TAPE40 = 3, PACK_ERR_SYSCALL = 4, PACK_ERR_OPCODE = 5, PACK_ERR_JALR = 6, PACK_ERR_ROM = 7,
  /* 8 = the lookups do not balance (prover.cu) */ PACK_ERR_SHAMT = 9, PACK_ERR_DIV0 = 10, PACK_ERR_MEMADDR = 11, PACK_ERR_MEMSIGN = 12,
  PACK_ERR_MEMVAL = 13, PACK_ERR_RAMROWS = 14,
};
BB_HD const char* pack_err_text(u32 code) {
  switch (code) {
    case PACK_ERR_PC: return "pc does not fit 30 bits";
    case PACK_ERR_REG40: return "register value exceeds 40 bits";
    case PACK_ERR_TAPE40: return "input tape value exceeds 40 bits";
    case PACK_ERR_SYSCALL: return "syscall other than EXIT/READ/WRITE/POSEIDON2";
    case PACK_ERR_OPCODE: return "opcode is not constrained";
    case PACK_ERR_JALR: return "jalr base register does not fit 30 bits";
    case PACK_ERR_ROM: return "pc is outside the program";
    case PACK_ERR_SHAMT: return "immediate shift amount above 63";
    case PACK_ERR_DIV0: return "division by zero";
    case PACK_ERR_MEMADDR: return "memory address outside [0, 2^30)";
    case PACK_ERR_MEMSIGN: return "lb / lh of a negative value (sign extension to 64 bits exceeds the 40-bit register model)";
    case PACK_ERR_MEMVAL: return "loaded value differs from the memory the AIR tracks (written by a syscall, or a data segment)";
    case PACK_ERR_RAMROWS: return "more RAM words touched than trace rows: use a larger log_n";
    default: return "?";
  }
}

BB_HD int pack_sext(u32 v, int bits) { const int sh = 32 - bits; return ((int)(v << sh)) >> sh; }
BB_HD u32 pack_inv(u32 canon) { return canon ? bb_from_mont(bb_inv(bb_to_mont(canon))) : 0u; }

// rg = PRE-state registers (rg[0] ignored), w = instruction word, read_val = post-state r10 (READ rows) / the loaded value (load rows).
// Wr: void operator()(int column, u32 canonical_value).  Returns a PackErr.
template <class Wr>
BB_HD u32 expand_row_v2(u64 i, u64 T, const u64 (&rg)[16], u64 pc, u32 w, u64 read_val, Wr& W) {
  const u64 LIMB = (1u << 20) - 1, M40 = (1ull << 40) - 1;
  const bool live = i < T;
  u32 err = PACK_OK;
  if (pc + 4 >= (1u << 30)) err = PACK_ERR_PC;
#pragma unroll
  for (int k = 1; k < 16; k++) if (rg[k] >> 40) err = PACK_ERR_REG40;

  W(ZKIR_COL_CLK, (u32)((live ? i : T) % BB_P));
  W(ZKIR_COL_PC, (u32)pc);
#pragma unroll
  for (int k = 1; k < 16; k++) { W(ZKIR_COL_R1_LO + 2 * (k - 1), (u32)(rg[k] & LIMB)); W(ZKIR_COL_R1_LO + 2 * (k - 1) + 1, (u32)((rg[k] >> 20) & LIMB)); }

  u32 rd = 0, rs1 = 0, rs2 = 0, writes = 0;
  u32 sel = 0xffffffffu;  // column of the opcode selector, if any
  u32 neg = 0, is_exit = 0, is_read = 0, is_write = 0, is_pos2 = 0;
  u64 av = 0, bv = 0, vv = 0, rc = 0;   // operands, written value, range-checked 40-bit quantity (-> chunks)
  long long imm = 0;
  bool has_imm = false;
  u32 carry0 = 0, carry1 = 0, taken = 0;
  u32 ch[4] = {0, 0, 0, 0};
  bool chunks_from_rc = false;
#ifdef ZKIR_PROFILE_FULL
  u64 fx = 0, fy = 0, fr = 0;            // multiplier block operands: X * Y + R = P
  bool f_mul = false, f_cmps = false, f_bit = false, f_shift = false, f_right = false, f_sra = false, f_div = false;
  u32 sa = 0, sb = 0, lts = 0, shamt = 0, shw = 0, zflag = 0;
  u64 fill = 0;
#endif
  if (live) {
    const u32 op = w & 0x7F;
    const u32 fa = (w >> 7) & 0xF, fb = (w >> 11) & 0xF, fc = (w >> 15) & 0xF;
    auto R = [&](u32 k) -> u64 {  // register read without dynamic indexing of the local array
      u64 v = 0;
#pragma unroll
      for (int j = 1; j < 16; j++) if (k == (u32)j) v = rg[j];
      return v;
    };
    const int imm17 = pack_sext((w >> 15) & 0x1FFFF, 17);
    if (op == 0x00 || op == 0x01) {                     // ADD / SUB
      rd = fa; rs1 = fb; rs2 = fc; writes = 1;
      av = R(rs1); bv = R(rs2);
      sel = op == 0 ? ZKIR_COL_S_ADD : ZKIR_COL_S_SUB;
    } else if (op == 0x08) {                            // ADDI
      rd = fa; rs1 = fb; imm = imm17; has_imm = true; writes = 1;
      av = R(rs1); bv = (u64)imm & M40;
      sel = ZKIR_COL_S_ADDI;
    } else if (op == 0x20 || op == 0x21) {              // SLTU / SGEU
      rd = fa; rs1 = fb; rs2 = fc; writes = 1; neg = op & 1;
      av = R(rs1); bv = R(rs2);
      sel = ZKIR_COL_S_SLTU;
    } else if (op == 0x24 || op == 0x25) {              // SEQ / SNE
      rd = fa; rs1 = fb; rs2 = fc; writes = 1; neg = op & 1;
      av = R(rs1); bv = R(rs2);
      sel = ZKIR_COL_S_SEQ;
    } else if (op == 0x26 || op == 0x27 || op == 0x28) { // CMOV / CMOVZ / CMOVNZ
      rd = fa; rs1 = fb; rs2 = fc; neg = op == 0x27;
      av = R(rs1); bv = R(rs2);
      sel = op == 0x28 ? ZKIR_COL_S_CMOVNZ : ZKIR_COL_S_CMOV;
    } else if (op == 0x40 || op == 0x41) {              // BEQ / BNE, B-type: rs1 bits 10:7, rs2 bits 14:11 (encoder.rs:132-140)
      rs1 = fa; rs2 = fb; imm = imm17; has_imm = true; neg = op & 1;
      av = R(rs1); bv = R(rs2);
      sel = ZKIR_COL_S_BEQ;
    } else if (op == 0x44 || op == 0x45) {              // BLTU / BGEU
      rs1 = fa; rs2 = fb; imm = imm17; has_imm = true; neg = op & 1;
      av = R(rs1); bv = R(rs2);
      sel = ZKIR_COL_S_BLTU;
    } else if (op == 0x48) {                            // JAL
      rd = fa; imm = pack_sext((w >> 11) & 0x1FFFFF, 21); has_imm = true; writes = 1;
      vv = pc + 4; rc = vv; chunks_from_rc = true;
      sel = ZKIR_COL_S_JAL;
    } else if (op == 0x49) {                            // JALR: target = (rs1 + imm) & ~1, link = pc + 4
      rd = fa; rs1 = fb; imm = imm17; has_imm = true; writes = 1;
      av = R(rs1);
      vv = pc + 4;
      if (av >> 30) err = PACK_ERR_JALR;
      ch[0] = (u32)(vv & 1023); ch[1] = (u32)((vv >> 10) & 1023); ch[2] = (u32)((vv >> 20) & 1023); ch[3] = (u32)((av >> 20) & 1023);
      taken = (u32)((av + (u64)(long long)imm) & 1);
      sel = ZKIR_COL_S_JALR;
    } else if (op == 0x50) {                            // ECALL: number in r10 (syscall.rs:94-149)
      const u64 num = rg[10];
      if (num == 0) is_exit = 1;
      else if (num == 1) {                              // READ: the value is the post-state r10
        is_read = 1; rd = 10; writes = 1;
        vv = read_val; rc = vv; chunks_from_rc = true;
        if (vv >> 40) err = PACK_ERR_TAPE40;
      } else if (num == 2) { is_write = 1; vv = rg[11]; }   // the written word goes to the I/O bus through v (no register is written: rd = r0)
      else if (num == 4) { is_pos2 = 1; rd = 10; writes = 1; vv = 0; }
      else err = PACK_ERR_SYSCALL;
    } else if (op == 0x51) {                            // EBREAK
      sel = ZKIR_COL_S_EBREAK;
#ifdef ZKIR_PROFILE_FULL
    } else if (op == 0x02 || op == 0x03) {              // MUL / MULH: low / high 40 bits of the 80-bit product
      rd = fa; rs1 = fb; rs2 = fc; writes = 1; neg = op & 1;
      av = R(rs1); bv = R(rs2); fx = av; fy = bv; f_mul = true;
      sel = ZKIR_COL_S_MUL;
    } else if (op >= 0x04 && op <= 0x07) {              // DIVU / REMU / DIV / REM (a provable value is below 2^40: the signed forms agree)
      rd = fa; rs1 = fb; rs2 = fc; writes = 1; neg = op & 1;
      av = R(rs1); bv = R(rs2);
      if (bv == 0) { err = PACK_ERR_DIV0; bv = 1; }
      fx = av / bv; fy = bv; fr = av % bv; f_mul = true; f_div = true;
      vv = neg ? fr : fx;
      sel = op < 0x06 ? ZKIR_COL_S_DIVU : ZKIR_COL_S_DIV;
    } else if (op >= 0x10 && op <= 0x15) {              // AND OR XOR / ANDI ORI XORI
      rd = fa; rs1 = fb; writes = 1;
      av = R(rs1);
      if (op >= 0x13) { imm = imm17; has_imm = true; bv = (u64)imm & M40; } else { rs2 = fc; bv = R(rs2); }
      fx = av; fy = bv; f_bit = true;
      const u32 k = (op - 0x10) % 3;
      vv = k == 0 ? (av & bv) : k == 1 ? (av | bv) : (av ^ bv);
      sel = ZKIR_COL_S_AND + (op - 0x10);
    } else if (op >= 0x18 && op <= 0x1D) {              // SLL SRL SRA / SLLI SRLI SRAI
      rd = fa; rs1 = fb; writes = 1;
      av = R(rs1);
      if (op >= 0x1B) { imm = (w >> 15) & 0xFF; has_imm = true; bv = (u64)imm; if (imm > 63) err = PACK_ERR_SHAMT; } else { rs2 = fc; bv = R(rs2); }
      f_mul = true; f_shift = true; f_right = (op - 0x18) % 3 != 0; f_sra = (op - 0x18) % 3 == 2;
      shamt = (u32)(bv & 63); shw = (u32)((bv & 1023) >> 6);
      fx = av; fr = bv;
      if (!f_right) fy = shamt < 40 ? 1ull << shamt : 0;
      else {
        fy = (shamt >= 1 && shamt <= 40) ? 1ull << (40 - shamt) : 0;
        zflag = shamt == 0;
        fill = shamt < 40 ? (((1ull << shamt) - 1) << (40 - shamt)) & M40 : M40;
      }
      if (f_sra) sa = (u32)((av >> 39) & 1);
      sel = ZKIR_COL_S_SLL + (op - 0x18);
    } else if (op == 0x22 || op == 0x23) {              // SLT / SGE
      rd = fa; rs1 = fb; rs2 = fc; writes = 1; neg = op & 1;
      av = R(rs1); bv = R(rs2); f_cmps = true;
      sel = ZKIR_COL_S_SLT;
    } else if (op == 0x42 || op == 0x43) {              // BLT / BGE
      rs1 = fa; rs2 = fb; imm = imm17; has_imm = true; neg = op & 1;
      av = R(rs1); bv = R(rs2); f_cmps = true;
      sel = ZKIR_COL_S_BLT;
    } else if ((op >= 0x30 && op <= 0x35) || (op >= 0x38 && op <= 0x3B)) {   // loads (I-type) / stores (S-type: base in bits 10:7, source in 14:11)
      imm = imm17; has_imm = true;
      if (op < 0x38) { rd = fa; rs1 = fb; writes = 1; vv = read_val; if (vv >> 40) err = PACK_ERR_REG40; sel = ZKIR_COL_S_LB + (op - 0x30); }
      else { rs1 = fa; rs2 = fb; bv = R(rs2); sel = ZKIR_COL_S_SB + (op - 0x38); }
      av = R(rs1);
      const u64 ea = av + (u64)imm;                     // execute.rs:478: rs1 + sign-extended offset
      if (ea >> 30) err = PACK_ERR_MEMADDR;
      rc = ea & M40; chunks_from_rc = true;
      const u64 i_lo = (u64)imm & LIMB, i_hi = imm < 0 ? LIMB : 0;
      const u64 k0 = ((av & LIMB) + i_lo) >> 20;
      carry0 = (u32)k0; carry1 = (u32)((((av >> 20) & LIMB) + i_hi + k0) >> 20);
#endif
    } else {
      err = PACK_ERR_OPCODE;
    }
    const u64 a_lo = av & LIMB, a_hi = (av >> 20) & LIMB, b_lo = bv & LIMB, b_hi = (bv >> 20) & LIMB;
    if (op == 0x00 || op == 0x08) {
      vv = (av + bv) & M40; rc = vv; chunks_from_rc = true;
      const u64 k0 = (a_lo + b_lo) >> 20;
      carry0 = (u32)k0; carry1 = (u32)((a_hi + b_hi + k0) >> 20);
#ifdef ZKIR_PROFILE_FULL
    } else if (f_cmps) {                                // signed compare at bit 39: (a <s b) = (a <u b) xor sign(a) xor sign(b)
      rc = (av - bv) & M40; chunks_from_rc = true;
      const u64 k0 = a_lo < b_lo;
      carry0 = (u32)k0; carry1 = (u32)(a_hi < b_hi + k0);
      fx = av; fy = bv;
      sa = (u32)((av >> 39) & 1); sb = (u32)((bv >> 39) & 1);
      lts = carry1 ^ sa ^ sb;
      if (op < 0x40) vv = lts ^ neg; else taken = lts ^ neg;
    } else if (f_div) {                                 // the chunks hold rem - divisor mod 2^40: the final borrow says rem < divisor
      rc = (fr - bv) & M40; chunks_from_rc = true;
      const u64 r_lo = fr & LIMB, r_hi = (fr >> 20) & LIMB;
      const u64 k0 = r_lo < b_lo;
      carry0 = (u32)k0; carry1 = (u32)(r_hi < b_hi + k0);
#endif
    } else if (op == 0x01 || op == 0x20 || op == 0x21 || op == 0x44 || op == 0x45) {
      rc = (av - bv) & M40; chunks_from_rc = true;    // SUB: the result; compares: a - b mod 2^40, final borrow = (a < b)
      const u64 k0 = a_lo < b_lo;
      carry0 = (u32)k0; carry1 = (u32)(a_hi < b_hi + k0);
      if (op == 0x01) vv = rc;
      else if (op < 0x40) vv = carry1 ^ neg;
      else taken = carry1 ^ neg;
    } else if (op == 0x24 || op == 0x25 || op == 0x40 || op == 0x41 || op == 0x26 || op == 0x27 || op == 0x28) {
      // is-zero gadget: EQ family on the limbs of a - b, CMOV family on the limbs of b
      const bool cm = op >= 0x26 && op <= 0x28;
      const u32 x_lo = cm ? (u32)b_lo : bb_sub((u32)a_lo, (u32)b_lo), x_hi = cm ? (u32)b_hi : bb_sub((u32)a_hi, (u32)b_hi);
      carry0 = x_lo != 0; carry1 = x_hi != 0;
      ch[0] = pack_inv(x_lo); ch[1] = pack_inv(x_hi);
      const u32 nz = carry0 | carry1;
      if (cm) {
        const u32 mv = neg ? !nz : nz;
        ch[3] = nz; ch[2] = mv; writes = mv; vv = av;
      } else {
        ch[2] = nz;
        const u32 eqx = neg ? nz : !nz;
        if (op < 0x40) vv = eqx; else taken = eqx;
      }
    }
  }
  if (chunks_from_rc) { ch[0] = (u32)(rc & 1023); ch[1] = (u32)((rc >> 10) & 1023); ch[2] = (u32)((rc >> 20) & 1023); ch[3] = (u32)((rc >> 30) & 1023); }
  u32 imm_lo = 0, imm_sign = 0;  // the high limb is the sign extension (0 or 2^20-1): not a column
  if (has_imm) {
    imm_lo = (u32)(((u64)imm & M40) & LIMB);
    imm_sign = imm < 0;
  }
  W(ZKIR_COL_IMM_LO, imm_lo); W(ZKIR_COL_IMM_SIGN, imm_sign);
#pragma unroll
  for (int c = ZKIR_COL_S_ADD; c <= ZKIR_COL_S_EBREAK; c++) W(c, sel == (u32)c);
#ifdef ZKIR_PROFILE_FULL
#pragma unroll
  for (int c = ZKIR_COL_S_MUL; c <= ZKIR_COL_S_SD; c++) W(c, sel == (u32)c);
#endif
  W(ZKIR_COL_NEG, neg);
  W(ZKIR_COL_IS_EXIT, is_exit); W(ZKIR_COL_IS_READ, is_read); W(ZKIR_COL_IS_WRITE, is_write); W(ZKIR_COL_IS_POS2, is_pos2);
#pragma unroll
  for (int k = 0; k < 3; k++) {  // register index = 4*h + l, two 4-way one-hots each (entry 3 implied)
    W(ZKIR_COL_RD_H0 + k, (rd >> 2) == (u32)k); W(ZKIR_COL_RD_L0 + k, (rd & 3u) == (u32)k);
    W(ZKIR_COL_RS1_H0 + k, (rs1 >> 2) == (u32)k); W(ZKIR_COL_RS1_L0 + k, (rs1 & 3u) == (u32)k);
    W(ZKIR_COL_RS2_H0 + k, (rs2 >> 2) == (u32)k); W(ZKIR_COL_RS2_L0 + k, (rs2 & 3u) == (u32)k);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) W(ZKIR_COL_RDW0 + k, ((rd >> 2) == (u32)k) ? writes : 0u);   // rdw[h] = rd_h[h] * write enable
  W(ZKIR_COL_A_LO, (u32)(av & LIMB)); W(ZKIR_COL_A_HI, (u32)((av >> 20) & LIMB));
  W(ZKIR_COL_B_LO, (u32)(bv & LIMB)); W(ZKIR_COL_B_HI, (u32)((bv >> 20) & LIMB));
  W(ZKIR_COL_V_LO, (u32)(vv & LIMB)); W(ZKIR_COL_V_HI, (u32)((vv >> 20) & LIMB));
#pragma unroll
  for (int k = 0; k < 4; k++) W(ZKIR_COL_CH0 + k, ch[k]);
  W(ZKIR_COL_CARRY0, carry0); W(ZKIR_COL_CARRY1, carry1);
  W(ZKIR_COL_TAKEN, taken);
#ifdef ZKIR_PROFILE_FULL
  {
    const bool dec = f_mul || f_cmps || f_bit;
    u32 x[4], y[4], r[4], p[8], klo[5], khi[5];
#pragma unroll
    for (int k = 0; k < 4; k++) { x[k] = dec ? (u32)((fx >> (10 * k)) & 1023) : 0u; y[k] = dec ? (u32)((fy >> (10 * k)) & 1023) : 0u; r[k] = (f_div || f_shift) ? (u32)((fr >> (10 * k)) & 1023) : 0u; }
    u64 c = 0;   // X * Y + R column by column in base 2^10 (only the DIV family adds R: the shifts keep rs2 in the r cells)
#pragma unroll
    for (int k = 0; k < 8; k++) {
      u64 t = c;
      if (f_mul) {
#pragma unroll
        for (int i = 0; i < 4; i++) if (k - i >= 0 && k - i < 4) t += (u64)x[i] * y[k - i];
        if (f_div && k < 4) t += r[k];
      }
      p[k] = (u32)(t & 1023); c = t >> 10;
      if (k < 5) { klo[k] = (u32)(c & 1023); khi[k] = (u32)(c >> 10); }
    }
    if (f_mul && !f_div && !f_shift) { const u64 lo = (u64)p[0] | (u64)p[1] << 10 | (u64)p[2] << 20 | (u64)p[3] << 30, hi = (u64)p[4] | (u64)p[5] << 10 | (u64)p[6] << 20 | (u64)p[7] << 30; vv = neg ? hi : lo; }
    if (f_shift) {
      const u64 lo = (u64)p[0] | (u64)p[1] << 10 | (u64)p[2] << 20 | (u64)p[3] << 30, hi = (u64)p[4] | (u64)p[5] << 10 | (u64)p[6] << 20 | (u64)p[7] << 30;
      vv = !f_right ? lo : hi + (zflag ? av : 0) + ((f_sra && sa) ? fill : 0);
    }
    if (f_mul || f_shift) { W(ZKIR_COL_V_LO, (u32)(vv & LIMB)); W(ZKIR_COL_V_HI, (u32)((vv >> 20) & LIMB)); }   // the product is known only now
#pragma unroll
    for (int k = 0; k < 4; k++) { W(ZKIR_COL_X0 + k, x[k]); W(ZKIR_COL_Y0 + k, y[k]); W(ZKIR_COL_R0 + k, r[k]); }
#pragma unroll
    for (int k = 0; k < 8; k++) W(ZKIR_COL_P0 + k, p[k]);
#pragma unroll
    for (int k = 0; k < 5; k++) { W(ZKIR_COL_K0_LO + k, klo[k]); W(ZKIR_COL_K0_HI + k, khi[k]); }
    W(ZKIR_COL_SIGN_A, sa); W(ZKIR_COL_SIGN_B, sb); W(ZKIR_COL_SIGN_XOR, sa ^ sb); W(ZKIR_COL_LT_SIGNED, lts);
    W(ZKIR_COL_SHAMT, shamt); W(ZKIR_COL_SH_W, shw); W(ZKIR_COL_SH_ZERO, zflag);
    W(ZKIR_COL_FILL_LO, (u32)(fill & LIMB)); W(ZKIR_COL_FILL_HI, (u32)((fill >> 20) & LIMB));
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 xl = f_bit ? x[k] & 31 : 0u, xh = f_bit ? x[k] >> 5 : 0u, yl = f_bit ? y[k] & 31 : 0u, yh = f_bit ? y[k] >> 5 : 0u;
      W(ZKIR_COL_XL0 + k, xl); W(ZKIR_COL_XH0 + k, xh); W(ZKIR_COL_YL0 + k, yl); W(ZKIR_COL_YH0 + k, yh);
      W(ZKIR_COL_ZL0 + k, xl & yl); W(ZKIR_COL_ZH0 + k, xh & yh);
    }
  }
#endif
  return err;
}

#ifdef ZKIR_PROFILE_FULL
// ---- full profile: the memory cells of a load / store row and the lookup multiplicities of a row, shared by the host packer and the
// device converter like the row function above (docs/PROVER_SPEC.md sections 3.7, 3.8)
struct MemAccess { bool is_ld, is_st; u32 op, width, off; u64 ea; };
// is `w` a load / store?  its width and effective address rs1 + sign-extended offset (execute.rs:478)
BB_HD bool mem_decode(u32 w, const u64 (&rg)[16], MemAccess& m) {
  m.op = w & 0x7F;
  m.is_ld = m.op >= 0x30 && m.op <= 0x35; m.is_st = m.op >= 0x38 && m.op <= 0x3B;
  if (!m.is_ld && !m.is_st) return false;
  const u32 ldw = m.op - 0x30, stw = m.op - 0x38;
  m.width = m.is_ld ? (ldw < 2 ? 1u : ldw < 4 ? 2u : ldw == 4 ? 4u : 8u) : (1u << stw);
  const u32 base = m.is_ld ? (w >> 11) & 15 : (w >> 7) & 15;
  u64 bv = 0;
#pragma unroll
  for (int j = 1; j < 16; j++) if (base == (u32)j) bv = rg[j];
  m.ea = bv + (u64)(long long)pack_sext((w >> 15) & 0x1FFFF, 17);
  m.off = (u32)(m.ea & 7);
  return true;
}
// cells of the memory row `i`: the word before (old_word, little endian; prev_ts = timestamp of its last access, 0 = initial) and after
// the access, offset one-hot, register-side bytes, timestamp distance.  *loaded = the value a load returns, *new_word = the word afterwards.
template <class Wr>
BB_HD u32 expand_mem_cells(u64 i, u32 w, const u64 (&rg)[16], const MemAccess& m, u64 old_word, u64 prev_ts, Wr& W, u64* loaded, u64* new_word) {
  u32 err = PACK_OK;
  if (m.ea >> 30) err = PACK_ERR_MEMADDR;
#pragma unroll
  for (int k = 0; k < 8; k++) W(ZKIR_COL_OFF0 + k, m.off == (u32)k);
  W(ZKIR_COL_MW, (u32)((m.ea & 1023) >> 3));
  u64 gval = 0, nw = old_word;
  if (m.is_ld) {
    const u64 sh = old_word >> (8 * m.off);
    gval = m.width == 8 ? sh : sh & ((1ull << (8 * m.width)) - 1);
    if ((m.op == 0x30 && (gval & 0x80)) || (m.op == 0x32 && (gval & 0x8000))) err = PACK_ERR_MEMSIGN;   // lb / lh of a negative value
    if (gval >> 40) err = PACK_ERR_REG40;
  } else {
    const u32 srcr = (w >> 11) & 15;
#pragma unroll
    for (int j = 1; j < 16; j++) if (srcr == (u32)j) gval = rg[j];
    if (gval >> 40) err = PACK_ERR_REG40;
    const u64 mask = m.width == 8 ? ~0ull : ((1ull << (8 * m.width)) - 1) << (8 * m.off);
    nw = (old_word & ~mask) | ((gval << (8 * m.off)) & mask);   // the low `width` bytes of rs2 (bytes 5..7 of an SD are zero: gval < 2^40)
  }
#pragma unroll
  for (int k = 0; k < 8; k++) { W(ZKIR_COL_OB0 + k, (u32)((old_word >> (8 * k)) & 255)); W(ZKIR_COL_NB0 + k, (u32)((nw >> (8 * k)) & 255)); }
#pragma unroll
  for (int k = 0; k < 5; k++) W(ZKIR_COL_GB0 + k, (u32)((gval >> (8 * k)) & 255));
  W(ZKIR_COL_NIB_LO, (u32)((gval >> 16) & 15)); W(ZKIR_COL_NIB_HI, (u32)((gval >> 20) & 15));
  W(ZKIR_COL_PREV_TS, (u32)(prev_ts % BB_P));
  const u64 dist = i - prev_ts;   // clk - prev_ts (prev_ts = clk' + 1 of the previous access, or 0)
#pragma unroll
  for (int k = 0; k < 3; k++) W(ZKIR_COL_TD0 + k, (u32)((dist >> (10 * k)) & 1023));
  *loaded = gval; *new_word = nw;
  return err;
}
// LogUp multiplicities a live row adds to the tables (tools/gen_air.py: fractions), from its opcode and its own cells:
// cell(column) -> value, add(table, index).  The range lookups of ch0..3 (row_range_checked) and the ROM row are the callers'.
enum MultTable : int { MT_RNG = 0, MT_AND, MT_POW, MT_B8, MT_B4, MT_B7 };
template <class Cell, class Add>
BB_HD void full_row_multiplicities(u32 w, Cell cell, Add add) {
  const u32 op = w & 0x7F;
  const bool mulf = op == 0x02 || op == 0x03, divf = op >= 0x04 && op <= 0x07, shf = op >= 0x18 && op <= 0x1D, cmps = op == 0x22 || op == 0x23 || op == 0x42 || op == 0x43;
  const bool bitf = op >= 0x10 && op <= 0x15, sraf = op == 0x1A || op == 0x1D, right = shf && (op - 0x18) % 3 != 0;
  const bool mul_on = mulf || divf || shf, is_ld = op >= 0x30 && op <= 0x35, is_st = op >= 0x38 && op <= 0x3B;
  if (mul_on || cmps || bitf) for (int k = 0; k < 4; k++) { add(MT_RNG, cell(ZKIR_COL_X0 + k)); add(MT_RNG, cell(ZKIR_COL_Y0 + k)); }
  if (divf || shf) for (int k = 0; k < 4; k++) add(MT_RNG, cell(ZKIR_COL_R0 + k));
  if (mul_on) {
    for (int k = 0; k < 8; k++) add(MT_RNG, cell(ZKIR_COL_P0 + k));
    for (int k = 0; k < 5; k++) { add(MT_RNG, cell(ZKIR_COL_K0_LO + k)); add(MT_RNG, cell(ZKIR_COL_K0_HI + k)); }
  }
  if (cmps || sraf) add(MT_RNG, 2 * (cell(ZKIR_COL_X0 + 3) - 512 * cell(ZKIR_COL_SIGN_A)));
  if (cmps) add(MT_RNG, 2 * (cell(ZKIR_COL_Y0 + 3) - 512 * cell(ZKIR_COL_SIGN_B)));
  if (shf) {
    add(MT_RNG, cell(ZKIR_COL_SH_W)); add(MT_RNG, 64 * cell(ZKIR_COL_SH_W));
    add(MT_POW, (right ? 64u : 0u) + cell(ZKIR_COL_SHAMT));
  }
  if (bitf) for (int k = 0; k < 4; k++) {
    add(MT_AND, cell(ZKIR_COL_XL0 + k) + 32 * cell(ZKIR_COL_YL0 + k));
    add(MT_AND, cell(ZKIR_COL_XH0 + k) + 32 * cell(ZKIR_COL_YH0 + k));
  }
  if (is_ld || is_st) {
    if (is_st) for (int k = 0; k < 5; k++) add(MT_B8, cell(ZKIR_COL_GB0 + k));
    add(MT_B4, cell(ZKIR_COL_NIB_LO)); add(MT_B4, cell(ZKIR_COL_NIB_HI));
    add(MT_B7, cell(ZKIR_COL_MW));
    if (op == 0x30) add(MT_B7, cell(ZKIR_COL_GB0));
    if (op == 0x32) add(MT_B7, cell(ZKIR_COL_GB0 + 1));
    for (int k = 0; k < 3; k++) add(MT_RNG, cell(ZKIR_COL_TD0 + k));
  }
}
#endif  // ZKIR_PROFILE_FULL

// does this row's chunk quadruple go to the range table?  (tools/gen_air.py: rc_on)
BB_HD bool row_range_checked(u32 w, u64 r10) {
  const u32 op = w & 0x7F;
#ifdef ZKIR_PROFILE_FULL
  if (op == 0x22 || op == 0x23 || op == 0x42 || op == 0x43 || (op >= 0x04 && op <= 0x07)) return true;   // signed compares, DIV family
  if ((op >= 0x30 && op <= 0x35) || (op >= 0x38 && op <= 0x3B)) return true;                              // loads / stores: the effective address
#endif
  return op == 0x00 || op == 0x01 || op == 0x08 || op == 0x20 || op == 0x21 || op == 0x44 || op == 0x45 || op == 0x48 || op == 0x49 ||
         (op == 0x50 && r10 == 1);
}

}  // namespace zkir
