// Public columns of the AIR (docs/PROVER_SPEC.md sections 3.3 and 3.7): tables every party can compute from the program alone.
// They are never committed: the prover extends them once per (program, shape), the verifier evaluates them at zeta itself.
//
// Both profiles: range table 0..1023 (zkir-spec/src/config.rs:76-80: 10-bit chunks), the program ROM as (pc, decoded word, imm)
// per instruction (zkir-assembler/src/encoder.rs:98-151).  Full profile only: the 5-bit x 5-bit AND table on the 1024 range rows
// and the power table of the shifts on rows 0..127 (zkir-spec/src/value.rs:658-691: shifts of 40 and more give 0 / the sign fill);
// the 8 / 4 / 7-bit range tables; the initial memory image as aligned 8-byte words: zeros below the load address 0x1000
// (zkir-runtime/src/memory.rs:297-309: uninitialised reads give 0), then the code words, little endian (zkir-runtime/src/vm.rs:138-170).
#include <stdint.h>
#include <stddef.h>
#include "zkir_b200.h"
#include "../air_profiles_generated.h"

extern "C" {
uint64_t zkir_image_words(size_t n_code);

// One row of the public columns: out[0 .. pub width of the profile).  Rows past the tables and the program hold the default row
// (i = ~0): zeros, decoded word 127 (no instruction has opcode 127), power-table copy of row 0.
void zkir_public_row(uint32_t width, uint64_t i, const uint32_t* code, size_t n_code, uint32_t* out) {
  const bool full = width == ZKIR_PROFILE_FULL_WIDTH;
  out[0] = i < 1024 ? (uint32_t)i : 0u;                                  // p_t
  if (i < n_code) { out[1] = 0x1000u + 4u * (uint32_t)i; zkir_rom_entry(code[i], &out[2], &out[3]); }
  else { out[1] = 0; out[2] = 127; out[3] = 0; }                         // p_pc, p_dec, p_imm
  if (!full) return;
  const uint32_t t = i < 1024 ? (uint32_t)i : 0u;
  out[4] = t & 31; out[5] = t >> 5; out[6] = (t & 31) & (t >> 5);        // p_ax, p_ay, p_az
  const uint64_t LIMB = (1u << 20) - 1, M40 = (1ull << 40) - 1;
  uint64_t key = 0, mul = 1, fill = 0; uint32_t zf = 0;                  // row 0: left shift by 0
  if (i < 64) { key = i; mul = i < 40 ? 1ull << i : 0; }
  else if (i < 128) {
    const uint64_t s = i - 64;
    key = 1024 + s; mul = (s >= 1 && s <= 40) ? 1ull << (40 - s) : 0; zf = s == 0;
    fill = s < 40 ? (((1ull << s) - 1) << (40 - s)) & M40 : M40;
  }
  out[7] = (uint32_t)key; out[8] = (uint32_t)(mul & LIMB); out[9] = (uint32_t)(mul >> 20); out[10] = zf;
  out[11] = (uint32_t)(fill & LIMB); out[12] = (uint32_t)(fill >> 20);   // p_key, p_mlo, p_mhi, p_zf, p_glo, p_ghi
  out[13] = i < 256 ? (uint32_t)i : 0u; out[14] = i < 16 ? (uint32_t)i : 0u; out[15] = i < 128 ? (uint32_t)i : 0u;   // p_b8, p_b4, p_b7
  const uint64_t n_img = zkir_image_words(n_code);
  const bool img = i < n_img;
  out[16] = img; out[17] = img ? (uint32_t)i : 0u;                         // p_img_on, p_img_a
  for (uint32_t k = 0; k < 8; k++) {                                       // p_img0..7: byte 8 i + k of the image
    const uint64_t addr = 8 * i + k;
    uint32_t b = 0;
    if (img && addr >= 0x1000 && (addr - 0x1000) / 4 < n_code) b = (code[(addr - 0x1000) / 4] >> (8 * (addr & 3))) & 255u;
    out[18 + k] = b;
  }
  out[26] = i == 0 ? (uint32_t)n_img : 0u;                                 // p_ram0
}

// aligned 8-byte words of the memory image [0, 0x1000 + 4 n_code): every word index from this on is RAM (starts as zero)
uint64_t zkir_image_words(size_t n_code) { return (0x1000 + 4 * (uint64_t)n_code + 7) / 8; }

// rows [0, zkir_public_rows) can differ from the default row
uint64_t zkir_public_rows(uint32_t width, size_t n_code) { (void)width; return n_code > 1024 ? n_code : 1024; }

// column-major [pub width][N]
void zkir_public_columns(uint32_t width, uint32_t log_n, const uint32_t* code, size_t n_code, uint32_t* cols) {
  const uint64_t N = 1ull << log_n;
  const uint32_t pw = width == ZKIR_PROFILE_FULL_WIDTH ? ZKIR_PROFILE_FULL_PUB : ZKIR_PROFILE_CORE_PUB;
  uint32_t row[ZKIR_PROFILE_MAX_PUB];
  for (uint64_t i = 0; i < N; i++) {
    zkir_public_row(width, i, code, n_code, row);
    for (uint32_t k = 0; k < pw; k++) cols[(size_t)k * N + i] = row[k];
  }
}

}  // extern "C"
