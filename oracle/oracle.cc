// zkir-b200 ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the trace->proof path that docs/PROVER_SPEC.md freezes.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// PARITY UNPINNED: /root/reference (seceq/zkir @ 82da87e9) contains no prover -- no NTT, no Poseidon2
// (zkir-runtime/src/crypto.rs:299-315 is a stub that returns Err), no Merkle/FRI, and Plonky3 is a
// commented-out, unpinned dependency (Cargo.toml:67-69).  There is therefore no reference output to pin
// this file against; what it follows is this repo's written spec.  The upstream (VM/trace) half *is*
// pinned by the reference's tests; that lives in zkir_b200/csrc/host/vm.cc and tests/test_vm_*.py.
//
// Deliberately dumb and structurally different from the CUDA path: canonical u32 values with 64-bit `%`
// reduction (the kernels use Montgomery form), textbook bit-reversal radix-2 NTT (the kernels use a
// multi-pass four-step), row-at-a-time hashing.  OpenMP only parallelises outer loops.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include "constants_generated.h"
// compiled once per AIR profile (docs/PROVER_SPEC.md section 3.7): liboracle.so = core, liboracle_full.so = full (-DZKIR_PROFILE_FULL)
#ifdef ZKIR_PROFILE_FULL
#include "air_generated_full.h"
#else
#include "air_generated.h"
#endif

typedef uint32_t u32;
typedef uint64_t u64;
static const u32 P = ZKIR_BB_P;

// ------------------------------------------------------------------ base field (canonical)
static inline u32 fadd(u32 a, u32 b) { u32 s = a + b; return s >= P ? s - P : s; }
static inline u32 fsub(u32 a, u32 b) { return a >= b ? a - b : a + P - b; }
static inline u32 fmul(u32 a, u32 b) { return (u32)(((u64)a * b) % P); }
static u32 fpow(u32 a, u64 e) { u32 r = 1; while (e) { if (e & 1) r = fmul(r, a); a = fmul(a, a); e >>= 1; } return r; }
static inline u32 finv(u32 a) { return fpow(a, P - 2); }

struct Fp {  // thin wrapper so air_generated.h can be instantiated with operators
  u32 v;
  Fp() : v(0) {}
  explicit Fp(u32 x) : v(x) {}
};
static inline Fp operator+(Fp a, Fp b) { return Fp(fadd(a.v, b.v)); }
static inline Fp operator-(Fp a, Fp b) { return Fp(fsub(a.v, b.v)); }
static inline Fp operator*(Fp a, Fp b) { return Fp(fmul(a.v, b.v)); }

// ------------------------------------------------------------------ extension F_p[X]/(X^4 - 11)
struct E4 { u32 c[4]; };
static inline E4 e4_zero() { E4 r = {{0, 0, 0, 0}}; return r; }
static inline E4 e4_from(u32 a) { E4 r = {{a, 0, 0, 0}}; return r; }
static inline E4 e4_add(E4 a, E4 b) { E4 r; for (int i = 0; i < 4; i++) r.c[i] = fadd(a.c[i], b.c[i]); return r; }
static inline E4 e4_sub(E4 a, E4 b) { E4 r; for (int i = 0; i < 4; i++) r.c[i] = fsub(a.c[i], b.c[i]); return r; }
static inline E4 e4_mulb(E4 a, u32 b) { E4 r; for (int i = 0; i < 4; i++) r.c[i] = fmul(a.c[i], b); return r; }
static E4 e4_mul(E4 a, E4 b) {
  // schoolbook product, then X^4 -> 11
  u32 t[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t[i + j] = fadd(t[i + j], fmul(a.c[i], b.c[j]));
  E4 r;
  for (int i = 0; i < 4; i++) r.c[i] = t[i];
  for (int i = 4; i < 7; i++) r.c[i - 4] = fadd(r.c[i - 4], fmul(ZKIR_EXT_W, t[i]));
  return r;
}
static E4 e4_pow(E4 a, u64 e) { E4 r = e4_from(1); while (e) { if (e & 1) r = e4_mul(r, a); a = e4_mul(a, a); e >>= 1; } return r; }
static E4 e4_inv(E4 a) {
  // dumbest correct way: a^(p^4 - 2) by square-and-multiply over a 124-bit exponent
  unsigned __int128 e = (unsigned __int128)P * P; e = e * P * P - 2;
  E4 r = e4_from(1);
  while (e) { if (e & 1) r = e4_mul(r, a); a = e4_mul(a, a); e >>= 1; }
  return r;
}
// Montgomery's trick: invert a whole vector with one field inversion (entries must be non-zero)
static void e4_batch_inv(E4* v, size_t n) {
  if (!n) return;
  std::vector<E4> pre(n);
  E4 acc = e4_from(1);
  for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = e4_mul(acc, v[i]); }
  E4 inv = e4_inv(acc);
  for (size_t i = n; i-- > 0;) { E4 t = e4_mul(inv, pre[i]); inv = e4_mul(inv, v[i]); v[i] = t; }
}
static inline bool e4_eq(E4 a, E4 b) { return !memcmp(a.c, b.c, 16); }

// ------------------------------------------------------------------ NTT (natural in, natural out)
static u32 root_of_unity(int logn) { return ZKIR_BB_ROOTS[logn]; }
static void ntt_inplace(u32* a, int logn, bool inverse) {
  size_t n = (size_t)1 << logn;
  for (size_t i = 0; i < n; i++) {  // bit reversal
    size_t j = 0;
    for (int b = 0; b < logn; b++) if (i >> b & 1) j |= (size_t)1 << (logn - 1 - b);
    if (j > i) std::swap(a[i], a[j]);
  }
  for (int s = 1; s <= logn; s++) {
    size_t m = (size_t)1 << s, h = m >> 1;
    u32 wm = root_of_unity(s);
    if (inverse) wm = finv(wm);
    for (size_t k = 0; k < n; k += m) {
      u32 w = 1;
      for (size_t j = 0; j < h; j++) {
        u32 t = fmul(w, a[k + j + h]), u = a[k + j];
        a[k + j] = fadd(u, t);
        a[k + j + h] = fsub(u, t);
        w = fmul(w, wm);
      }
    }
  }
  if (inverse) { u32 ninv = finv((u32)(n % P)); for (size_t i = 0; i < n; i++) a[i] = fmul(a[i], ninv); }
}
// evaluations of the degree<N polynomial with coefficient vector c on the coset shift*H_{M}, M = N<<logb
static void coset_eval(const u32* coef, size_t ncoef, u32* out, int logm, u32 shift) {
  size_t m = (size_t)1 << logm;
  u32 s = 1;
  for (size_t j = 0; j < m; j++) { out[j] = j < ncoef ? fmul(coef[j], s) : 0; if (j < ncoef) s = fmul(s, shift); }
  ntt_inplace(out, logm, false);
}

// ------------------------------------------------------------------ Poseidon2 (width 16, x^7, 8+13)
static inline u32 sbox(u32 x) { u32 x2 = fmul(x, x), x3 = fmul(x2, x), x4 = fmul(x2, x2); return fmul(x4, x3); }
static void m4(u32* x) {  // circ(2,3,1,1)
  u32 y[4];
  for (int j = 0; j < 4; j++) {
    u32 a = x[j], b = x[(j + 1) & 3], c = x[(j + 2) & 3], d = x[(j + 3) & 3];
    y[j] = fadd(fadd(fadd(a, a), fadd(fadd(b, b), b)), fadd(c, d));
  }
  memcpy(x, y, 16);
}
static void external_linear(u32* s) {
  for (int c = 0; c < 4; c++) m4(s + 4 * c);
  u32 sums[4];
  for (int k = 0; k < 4; k++) sums[k] = fadd(fadd(s[k], s[4 + k]), fadd(s[8 + k], s[12 + k]));
  for (int i = 0; i < 16; i++) s[i] = fadd(s[i], sums[i & 3]);
}
static void internal_linear(u32* s) {
  u32 sum = 0;
  for (int i = 0; i < 16; i++) sum = fadd(sum, s[i]);
  for (int i = 0; i < 16; i++) s[i] = fadd(fmul(s[i], ZKIR_P2_DIAG[i]), sum);
}
static void poseidon2(u32* s) {
  external_linear(s);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 16; i++) s[i] = sbox(fadd(s[i], ZKIR_P2_RC_EXT[r * 16 + i]));
    external_linear(s);
  }
  for (int r = 0; r < ZKIR_P2_RP; r++) {
    s[0] = sbox(fadd(s[0], ZKIR_P2_RC_INT[r]));
    internal_linear(s);
  }
  for (int r = 4; r < 8; r++) {
    for (int i = 0; i < 16; i++) s[i] = sbox(fadd(s[i], ZKIR_P2_RC_EXT[r * 16 + i]));
    external_linear(s);
  }
}
// overwrite-mode sponge, rate 8, digest 8 (no padding)
static void hash_elems(const u32* in, size_t n, u32* digest) {
  u32 st[16] = {0};
  for (size_t i = 0; i < n; i += 8) {
    size_t len = std::min<size_t>(8, n - i);
    for (size_t k = 0; k < len; k++) st[k] = in[i + k];
    poseidon2(st);
  }
  memcpy(digest, st, 32);
}
static void compress(const u32* l, const u32* r, u32* out) {
  u32 st[16];
  memcpy(st, l, 32); memcpy(st + 8, r, 32);
  poseidon2(st);
  memcpy(out, st, 32);
}
// tree layout: level 0 = leaves [n][8], then n/2 nodes, ... , root; total (2n-1)*8 words
static void merkle_build(u32* tree, size_t nleaves) {
  u32* lvl = tree;
  for (size_t n = nleaves; n > 1; n >>= 1) {
    u32* nxt = lvl + n * 8;
#pragma omp parallel for if (n > 256)
    for (size_t i = 0; i < n / 2; i++) compress(lvl + 16 * i, lvl + 16 * i + 8, nxt + 8 * i);
    lvl = nxt;
  }
}
static const u32* merkle_root(const u32* tree, size_t nleaves) { return tree + (2 * nleaves - 2) * 8; }
// digest of a long message for the transcript (docs/PROVER_SPEC.md section 2, `hash_tree`): chunks of 8 words are leaves
// (last one may be short), padded with all-zero digests to a power of two, Merkle root.  Log-depth instead of a serial sponge.
static void hash_tree(const u32* words, size_t n, u32* digest) {
  size_t chunks = (n + 7) / 8, leaves = 1;
  while (leaves < chunks) leaves <<= 1;
  std::vector<u32> tree((2 * leaves - 1) * 8, 0);
  for (size_t c = 0; c < chunks; c++) hash_elems(words + 8 * c, n - 8 * c < 8 ? n - 8 * c : 8, &tree[c * 8]);
  merkle_build(tree.data(), leaves);
  memcpy(digest, merkle_root(tree.data(), leaves), 32);
}
static void merkle_path(const u32* tree, size_t nleaves, size_t idx, u32* out) {
  const u32* lvl = tree;
  for (size_t n = nleaves; n > 1; n >>= 1) { memcpy(out, lvl + ((idx ^ 1) * 8), 32); out += 8; lvl += n * 8; idx >>= 1; }
}

// ------------------------------------------------------------------ duplex challenger (width 16, rate 8)
struct Chal {
  u32 st[16]; u32 in[8]; int nin; u32 out[8]; int nout;
  Chal() { memset(this, 0, sizeof(*this)); }
  void duplex() { for (int i = 0; i < nin; i++) st[i] = in[i]; nin = 0; poseidon2(st); memcpy(out, st, 32); nout = 8; }
  void observe(u32 x) { nout = 0; in[nin++] = x; if (nin == 8) duplex(); }
  void observe_n(const u32* x, size_t n) { for (size_t i = 0; i < n; i++) observe(x[i]); }
  u32 sample() { if (nin > 0 || nout == 0) duplex(); return out[--nout]; }
  E4 sample_ext() { E4 r; for (int i = 0; i < 4; i++) r.c[i] = sample(); return r; }
  u32 sample_bits(int b) { u32 v = sample(); return b >= 32 ? v : (v & ((1u << b) - 1)); }
};

// ------------------------------------------------------------------ AIR evaluation contexts (air_generated.h, AIR v2)
struct Xp { E4 v; };  // ext4 with operators, for the generated code
static inline Xp operator+(Xp a, Xp b) { Xp r; r.v = e4_add(a.v, b.v); return r; }
static inline Xp operator-(Xp a, Xp b) { Xp r; r.v = e4_sub(a.v, b.v); return r; }
static inline Xp operator*(Xp a, Xp b) { Xp r; r.v = e4_mul(a.v, b.v); return r; }
static inline Xp operator*(Xp a, Fp b) { Xp r; r.v = e4_mulb(a.v, b.v); return r; }

struct LookupChallenges { E4 z, th[ZKIR_AIR_NUM_THETA + 1], sio; };  // th[k] = theta^k; sio = sum of the public I/O transcript's fractions
static LookupChallenges make_challenges(E4 z, E4 theta) {
  LookupChallenges c; c.z = z; c.th[0] = e4_from(1); c.th[1] = theta;
  for (int k = 2; k <= ZKIR_AIR_NUM_THETA; k++) c.th[k] = e4_mul(c.th[k - 1], theta);
  c.sio = e4_zero();
  return c;
}
// public I/O transcript: n events of 4 words (clk, kind: 0 READ / 1 WRITE, value lo, value hi); its LogUp sum
//   S_io = sum_e 1 / (z - (3 + theta*clk + theta^2*kind + theta^3*lo + theta^4*hi))
static E4 e4_inv(E4 a);
static E4 io_sum(const LookupChallenges& c, const u32* io, size_t n_io) {
  E4 s = e4_zero();
  for (size_t e = 0; e < n_io; e++) {
    E4 fp = e4_from(3);
    for (int k = 0; k < 4; k++) fp = e4_add(fp, e4_mulb(c.th[k + 1], io[4 * e + k] % P));
    s = e4_add(s, e4_inv(e4_sub(c.z, fp)));
  }
  return s;
}

struct AirRowCtx {  // all constraints at one point: main/aux/public columns are column-major with the same stride
  typedef Fp F; typedef Xp X;
  const u32 *data, *aux, *pub; size_t stride, row, nxt; const u32* pv; const LookupChallenges* lc;
  Fp is_first, is_last, is_trans;
  E4 vals[ZKIR_AIR_NUM_CONSTRAINTS];
  Fp L(int i) const { return Fp(data[i * stride + row]); }
  Fp N(int i) const { return Fp(data[i * stride + nxt]); }
  Fp A(int i) const { return Fp(aux[i * stride + row]); }
  Fp AN(int i) const { return Fp(aux[i * stride + nxt]); }
  Fp P(int i) const { return Fp(pub[i * stride + row]); }
  Fp PV(int i) const { return Fp(pv[i]); }
  Fp K(u32 k) const { return Fp(k); }
  Xp z() const { Xp r; r.v = lc->z; return r; }
  Xp th(int k) const { Xp r; r.v = lc->th[k]; return r; }
  Xp xf(Fp a) const { Xp r; r.v = e4_from(a.v); return r; }
  Xp x4(Fp a, Fp b, Fp c, Fp d) const { Xp r; r.v.c[0] = a.v; r.v.c[1] = b.v; r.v.c[2] = c.v; r.v.c[3] = d.v; return r; }
  Xp sio() const { Xp r; r.v = lc->sio; return r; }
  void fence() const {}
  void emit(int idx, Fp v) { vals[idx] = e4_from(v.v); }
  void emit_x(int idx, Xp v) { vals[idx] = v.v; }
};
struct FracCtx {  // the LogUp fractions of one row (zkir_air_fractions): numerators and denominators
  typedef Fp F; typedef Xp X;
  const u32 *data, *pub; size_t stride, row; const LookupChallenges* lc;
  u32 num[ZKIR_AIR_NUM_FRACTIONS]; E4 den[ZKIR_AIR_NUM_FRACTIONS];
  Fp L(int i) const { return Fp(data[i * stride + row]); }
  Fp P(int i) const { return Fp(pub[i * stride + row]); }
  Fp K(u32 k) const { return Fp(k); }
  Xp z() const { Xp r; r.v = lc->z; return r; }
  Xp th(int k) const { Xp r; r.v = lc->th[k]; return r; }
  Xp xf(Fp a) const { Xp r; r.v = e4_from(a.v); return r; }
  void frac(int j, Fp n, Xp d) { num[j] = n.v; den[j] = d.v; }
};

// ---- public columns (docs/PROVER_SPEC.md section 3.3): range table and decoded program ROM, [4][N] column-major
static const u32 CODE_BASE = 0x1000;
static inline int32_t sext32(u32 v, int bits) { int sh = 32 - bits; return ((int32_t)(v << sh)) >> sh; }
// decoded word of the ROM: opcode | rd << 7 | rs1 << 11 | rs2 << 15 with the fields of the word's FORMAT (encoder.rs:98-151):
// R: rd rs1 rs2; I: rd rs1 imm17; shift-immediate: rd rs1 shamt; S/B: rs1 (bits 10:7) rs2 (bits 14:11) imm17; J: rd imm21; ECALL/EBREAK: none
static void rom_decode(u32 w, u32* dec, u32* imm) {
  const u32 op = w & 0x7F, f7 = (w >> 7) & 15, f11 = (w >> 11) & 15, f15 = (w >> 15) & 15;
  int64_t im = 0; u32 rd = 0, rs1 = 0, rs2 = 0;
  const bool itype = op == 0x08 || (op >= 0x13 && op <= 0x15) || (op >= 0x30 && op <= 0x35) || op == 0x49;
  if (itype) { rd = f7; rs1 = f11; im = sext32((w >> 15) & 0x1FFFF, 17); }
  else if (op >= 0x1B && op <= 0x1D) { rd = f7; rs1 = f11; im = (w >> 15) & 0xFF; }
  else if ((op >= 0x38 && op <= 0x3B) || (op >= 0x40 && op <= 0x45)) { rs1 = f7; rs2 = f11; im = sext32((w >> 15) & 0x1FFFF, 17); }
  else if (op == 0x48) { rd = f7; im = sext32((w >> 11) & 0x1FFFFF, 21); }
  else if (op == 0x50 || op == 0x51) {}
  else { rd = f7; rs1 = f11; rs2 = f15; }   // R-type (and anything undefined: it can never match a trace row)
  *dec = op | rd << 7 | rs1 << 11 | rs2 << 15;
  *imm = im < 0 ? (u32)(P + im) : (u32)im;
}
static void build_public_columns(u32 log_n, const u32* code, size_t n_code, u32* pub) {
  const size_t N = (size_t)1 << log_n;
  memset(pub, 0, ZKIR_AIR_PUB_WIDTH * N * sizeof(u32));
  for (size_t i = 0; i < N && i < ((size_t)1 << ZKIR_AIR_RANGE_BITS); i++) pub[0 * N + i] = (u32)i;
  for (size_t i = 0; i < N; i++) {
    if (i < n_code) { pub[1 * N + i] = CODE_BASE + 4 * (u32)i; rom_decode(code[i], &pub[2 * N + i], &pub[3 * N + i]); }
    else pub[2 * N + i] = 127;   // no instruction has opcode 127: an unused ROM row matches no trace row
  }
#ifdef ZKIR_PROFILE_FULL
  // docs/PROVER_SPEC.md section 3.7.  Columns 4..6: AND table, row t = x + 32 y (t < 1024) holds (x, y, x & y).  Columns 7..12: power
  // table (key, multiplier lo/hi limb, zero flag, fill lo/hi limb): rows 0..63 = left shift by s (key s, 2^s or 0 from 40 on),
  // rows 64..127 = right shift by s (key 1024 + s, 2^(40-s) for 1 <= s <= 40 else 0, flag s == 0, fill = top min(s, 40) bits of a
  // 40-bit word); every other row repeats row 0.
  for (size_t i = 0; i < N; i++) {
    if (i < 1024) { const u32 x = i & 31, y = (u32)i >> 5; pub[4 * N + i] = x; pub[5 * N + i] = y; pub[6 * N + i] = x & y; }
    u64 key = 0, mul = 1, fill = 0; u32 zf = 0;
    if (i >= 1 && i < 64) { key = i; mul = i < 40 ? (u64)1 << i : 0; }
    if (i >= 64 && i < 128) {
      const u64 sh = i - 64;
      key = 1024 + sh; zf = sh == 0; mul = (sh >= 1 && sh <= 40) ? (u64)1 << (40 - sh) : 0;
      for (u64 b = 0; b < sh && b < 40; b++) fill |= (u64)1 << (39 - b);
    }
    pub[7 * N + i] = (u32)key; pub[8 * N + i] = (u32)(mul & 0xFFFFF); pub[9 * N + i] = (u32)(mul >> 20); pub[10 * N + i] = zf;
    pub[11 * N + i] = (u32)(fill & 0xFFFFF); pub[12 * N + i] = (u32)(fill >> 20);
  }
  // docs/PROVER_SPEC.md section 3.8.  Columns 13..15: 8 / 4 / 7-bit range tables (row t holds t while t < 2^k, else 0).  Columns 16..25: the
  // memory image, one aligned 8-byte word per row: flag, word index, the eight bytes (zeros below 0x1000, then the code words, little
  // endian).  Column 26: the first RAM word index (= number of image words), on row 0 only.
  const size_t n_img = (CODE_BASE + 4 * n_code + 7) / 8;
  for (size_t i = 0; i < N; i++) {
    if (i < 256) pub[13 * N + i] = (u32)i;
    if (i < 16) pub[14 * N + i] = (u32)i;
    if (i < 128) pub[15 * N + i] = (u32)i;
    if (i < n_img) {
      pub[16 * N + i] = 1; pub[17 * N + i] = (u32)i;
      for (size_t k = 0; k < 8; k++) {
        const size_t addr = 8 * i + k;
        if (addr >= CODE_BASE && (addr - CODE_BASE) / 4 < n_code) pub[(18 + k) * N + i] = (code[(addr - CODE_BASE) / 4] >> (8 * (addr % 4))) & 0xFF;
      }
    }
  }
  pub[26 * N + 0] = (u32)n_img;
#endif
}
// aux columns [16][N] from the main trace, the public columns and the lookup challenges: helper k = sum of its two fractions,
// phi = running sum of all fractions of the earlier rows.  Returns false if the lookups do not balance (invalid witness).
static bool build_aux(u32 log_n, const u32* trace, const u32* pub, const LookupChallenges& lc, u32* aux) {
  const size_t N = (size_t)1 << log_n;
  const int NF = ZKIR_AIR_NUM_FRACTIONS, NH = ZKIR_AIR_NUM_HELPERS;
  static const int helper_of[NF] = ZKIR_AIR_FRAC_HELPER_INIT;
  std::vector<E4> den(N * NF), tot(N);
  std::vector<u32> num(N * NF);
#pragma omp parallel for
  for (size_t i = 0; i < N; i++) {
    FracCtx c; c.data = trace; c.pub = pub; c.stride = N; c.row = i; c.lc = &lc;
    zkir_air_fractions(c);
    for (int j = 0; j < NF; j++) { num[i * NF + j] = c.num[j]; den[i * NF + j] = c.den[j]; }
  }
  const size_t CH = 4096 * NF;
#pragma omp parallel for
  for (size_t c0 = 0; c0 < N * NF; c0 += CH) e4_batch_inv(&den[c0], std::min(CH, N * NF - c0));
#pragma omp parallel for
  for (size_t i = 0; i < N; i++) {
    E4 h[NH + 1];   // helper sums; h[NH] = the fractions the running sum adds itself
    for (int k = 0; k <= NH; k++) h[k] = e4_zero();
    for (int j = 0; j < NF; j++) h[helper_of[j]] = e4_add(h[helper_of[j]], e4_mulb(den[i * NF + j], num[i * NF + j]));
    E4 t = h[NH];
    for (int k = 0; k < NH; k++) { for (int q = 0; q < 4; q++) aux[(4 * k + q) * N + i] = h[k].c[q]; t = e4_add(t, h[k]); }
    tot[i] = t;
  }
  E4 phi = e4_zero();
  for (size_t i = 0; i < N; i++) { for (int q = 0; q < 4; q++) aux[(4 * NH + q) * N + i] = phi.c[q]; phi = e4_add(phi, tot[i]); }
  return e4_eq(phi, lc.sio);   // range and ROM fractions cancel, the I/O rows must add up to the public transcript's sum
}
// digest of the public I/O transcript the transcript absorbs: hash_tree over {n_io, clk_0, kind_0, lo_0, hi_0, ...}
static void hash_tree(const u32* words, size_t n, u32* digest);
static void io_digest(const u32* io, size_t n_io, u32* digest) {
  std::vector<u32> w(4 * n_io + 1);
  w[0] = (u32)n_io;
  for (size_t i = 0; i < 4 * n_io; i++) w[1 + i] = io[i] % P;
  hash_tree(w.data(), w.size(), digest);
}
// digest of the program the transcript absorbs: hash_tree over the 16-bit halves of the code words (each < p)
static void program_digest(const u32* code, size_t n_code, u32* digest) {
  std::vector<u32> halves(2 * n_code + 1);
  halves[0] = (u32)n_code;
  for (size_t i = 0; i < n_code; i++) { halves[1 + 2 * i] = code[i] & 0xFFFF; halves[2 + 2 * i] = code[i] >> 16; }
  hash_tree(halves.data(), halves.size(), digest);
}

struct Params { u32 log_blowup, num_queries, pow_bits, width, num_public; };
static const u32 PROOF_MAGIC = 0x5A4B5052u, PROOF_VERSION = 6u;
static const size_t AW = ZKIR_AIR_AUX_WIDTH, PW = ZKIR_AIR_PUB_WIDTH;

// FRI rounds (docs/PROVER_SPEC.md section 4.6): log_n / 3 rounds that fold by 8, then one that folds by 2^(log_n mod 3) if that is > 1
static size_t fri_rounds(u32 log_n) { return log_n / 3 + (log_n % 3 ? 1 : 0); }
static u32 fri_log_arity(u32 log_n, size_t t) { return t < log_n / 3 ? 3 : log_n % 3; }
static size_t proof_words(const Params& p, u32 log_n) {
  size_t lg = log_n + p.log_blowup, W = p.width, WA = W + AW, R = fri_rounds(log_n);
  size_t n = 8 + p.num_public + 24 + (2 * WA + 8) * 4 + R * 8 + 4 + 1;
  const size_t lr = (size_t)2 << p.log_blowup, depth = lg - (p.log_blowup + 1);   // rows per Merkle leaf of the LDE matrices, tree depth
  size_t perq = lr * (W + AW + 8) + 3 * depth * 8;
  size_t ll = lg;  // log2 of the layer length
  for (size_t t = 0; t < R; t++) { const u32 la = fri_log_arity(log_n, t); perq += (4u << la) + (ll - la) * 8; ll -= la; }
  return n + perq * p.num_queries;
}

// Quotient values on the LDE coset, natural order: out[4][M] (ext4 coefficient planes).  lde / aux / pub: the LDEs of the main,
// aux and public columns, column-major with stride M.
static void quotient_evals(const Params& p, u32 log_n, const u32* lde, const u32* aux, const u32* pub, const u32* pv,
                           const LookupChallenges& lc, E4 alpha, u32* out) {
  const int K = ZKIR_AIR_NUM_CONSTRAINTS;
  size_t N = (size_t)1 << log_n, B = (size_t)1 << p.log_blowup, M = N * B;
  u32 w = root_of_unity(log_n + p.log_blowup), g_inv = finv(root_of_unity(log_n));
  std::vector<E4> apow(K);  // apow[i] = alpha^(K-1-i)
  E4 a = e4_from(1);
  for (int i = K - 1; i >= 0; i--) { apow[i] = a; a = e4_mul(a, alpha); }
  std::vector<u32> xs(M);
  u32 x = ZKIR_BB_GEN;
  for (size_t i = 0; i < M; i++) { xs[i] = x; x = fmul(x, w); }
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < M; i++) {
    u32 xi = xs[i];
    u32 zh = fsub(fpow(xi, N), 1);                 // Z_H(x) = x^N - 1
    u32 zh_inv = finv(zh);
    AirRowCtx c;
    c.data = lde; c.aux = aux; c.pub = pub; c.stride = M; c.row = i; c.nxt = (i + B) % M; c.pv = pv; c.lc = &lc;
    c.is_first = Fp(fmul(zh, finv(fsub(xi, 1))));  // Z_H(x)/(x-1)
    c.is_last = Fp(fmul(zh, finv(fsub(xi, g_inv))));
    c.is_trans = Fp(fsub(xi, g_inv));
    zkir_air_eval(c);
    E4 acc = e4_zero();
    for (int k = 0; k < K; k++) acc = e4_add(acc, e4_mul(apow[k], c.vals[k]));
    acc = e4_mulb(acc, zh_inv);
    for (int k = 0; k < 4; k++) out[k * M + i] = acc.c[k];
  }
}

// Merkle commitment of an LDE matrix (columns [c0, c0 + nc) of `mat`, column stride M): leaf m = hash of the 2*B consecutive
// natural-order rows m*2B .. m*2B + 2B - 1 concatenated (docs/PROVER_SPEC.md section 4.1); returns the tree ((2*M/2B - 1) * 8 words)
static std::vector<u32> commit_matrix(const u32* mat, size_t M, size_t nc, u32 log_blowup) {
  const size_t lr = (size_t)2 << log_blowup, leaves = M / lr;
  std::vector<u32> tree((2 * leaves - 1) * 8);
#pragma omp parallel for
  for (size_t m = 0; m < leaves; m++) {
    std::vector<u32> buf(lr * nc);
    for (size_t r = 0; r < lr; r++) for (size_t k = 0; k < nc; k++) buf[r * nc + k] = mat[k * M + m * lr + r];
    hash_elems(buf.data(), buf.size(), &tree[m * 8]);
  }
  merkle_build(tree.data(), leaves);
  return tree;
}

struct Dump {  // optional intermediates for stage-by-stage parity tests
  u32 alpha[4], zeta[4], alpha_fri[4], lookup_z[4], lookup_theta[4];
  u32* aux;          // [16][N] or null
  u32* quotient;     // [4][M] or null
  u32* fri_input;    // [M][4] or null
  u32* lde;          // [W][M] or null
  u32* betas;        // [R][4] or null
};

static int prove(const Params& p, const u32* trace, u32 log_n, const u32* pv, const u32* code, size_t n_code, const u32* io, size_t n_io, u32* proof, Dump* dump) {
  const size_t W = p.width, WA = W + AW, N = (size_t)1 << log_n, B = (size_t)1 << p.log_blowup, M = N * B;
  const int lg = log_n + p.log_blowup;
  const u32 shift = ZKIR_BB_GEN;
  if (W != ZKIR_AIR_WIDTH || p.num_public != ZKIR_AIR_NUM_PUBLIC) return -1;
  if (log_n < ZKIR_AIR_RANGE_BITS || n_code > N) return -6;   // the range table and the ROM must fit the trace
  u32* out = proof;
  *out++ = PROOF_MAGIC; *out++ = PROOF_VERSION; *out++ = log_n; *out++ = p.width; *out++ = p.log_blowup;
  *out++ = p.num_queries; *out++ = p.pow_bits; *out++ = p.num_public;
  for (u32 i = 0; i < p.num_public; i++) *out++ = pv[i];

  // ---- 1. LDE of the main trace (committed) and of the public columns (known to the verifier, not committed)
  // coef / lde hold main columns [0, W) then aux columns [W, W + AW)
  std::vector<u32> coef(WA * N), lde(WA * M), pub(PW * N), publde(PW * M);
  build_public_columns(log_n, code, n_code, pub.data());
#pragma omp parallel for
  for (size_t k = 0; k < W + PW; k++) {
    if (k < W) {
      memcpy(&coef[k * N], trace + k * N, N * 4);
      ntt_inplace(&coef[k * N], log_n, true);
      coset_eval(&coef[k * N], N, &lde[k * M], lg, shift);
    } else {
      std::vector<u32> c(pub.begin() + (k - W) * N, pub.begin() + (k - W + 1) * N);
      ntt_inplace(c.data(), log_n, true);
      coset_eval(c.data(), N, &publde[(k - W) * M], lg, shift);
    }
  }
  const size_t LR = (size_t)2 << p.log_blowup, LEAVES = M / LR;
  std::vector<u32> ttree = commit_matrix(lde.data(), M, W, p.log_blowup);
  Chal ch;
  ch.observe(log_n); ch.observe(p.width); ch.observe((u32)AW); ch.observe(p.log_blowup); ch.observe(p.num_queries); ch.observe(p.pow_bits);
  ch.observe(p.num_public);
  ch.observe_n(pv, p.num_public);
  {
    u32 pd[8];
    program_digest(code, n_code, pd);
    ch.observe_n(pd, 8);
    io_digest(io, n_io, pd);
    ch.observe_n(pd, 8);
  }
  ch.observe_n(merkle_root(ttree.data(), LEAVES), 8);
  memcpy(out, merkle_root(ttree.data(), LEAVES), 32); out += 8;

  // ---- 1b. lookup challenges, aux columns (LogUp helpers and running sum), their LDE and commitment
  E4 lz = ch.sample_ext(), ltheta = ch.sample_ext();
  LookupChallenges lc = make_challenges(lz, ltheta);
  lc.sio = io_sum(lc, io, n_io);
  {
    std::vector<u32> aux(AW * N);
    if (!build_aux(log_n, trace, pub.data(), lc, aux.data())) return -7;   // lookups do not balance: invalid witness
    if (dump) { memcpy(dump->lookup_z, lz.c, 16); memcpy(dump->lookup_theta, ltheta.c, 16); if (dump->aux) memcpy(dump->aux, aux.data(), AW * N * 4); }
#pragma omp parallel for
    for (size_t k = 0; k < AW; k++) {
      memcpy(&coef[(W + k) * N], &aux[k * N], N * 4);
      ntt_inplace(&coef[(W + k) * N], log_n, true);
      coset_eval(&coef[(W + k) * N], N, &lde[(W + k) * M], lg, shift);
    }
  }
  if (dump && dump->lde) memcpy(dump->lde, lde.data(), WA * M * 4);
  std::vector<u32> atree = commit_matrix(lde.data() + W * M, M, AW, p.log_blowup);
  ch.observe_n(merkle_root(atree.data(), LEAVES), 8);
  memcpy(out, merkle_root(atree.data(), LEAVES), 32); out += 8;
  u32* quot_root_slot = out; out += 8;

  // ---- 2. quotient
  E4 alpha = ch.sample_ext();
  std::vector<u32> q(4 * M);
  quotient_evals(p, log_n, lde.data(), lde.data() + W * M, publde.data(), pv, lc, alpha, q.data());
  if (dump) { memcpy(dump->alpha, alpha.c, 16); if (dump->quotient) memcpy(dump->quotient, q.data(), 4 * M * 4); }
  // coefficients of Q: inverse NTT on the coset, then undo the shift
  std::vector<u32> qcoef(8 * N), qlde(8 * M);
  u32 sinv = finv(shift);
  for (int k = 0; k < 4; k++) {
    ntt_inplace(&q[k * M], lg, true);
    u32 s = 1;
    for (size_t j = 0; j < M; j++) { q[k * M + j] = fmul(q[k * M + j], s); s = fmul(s, sinv); }
    for (size_t j = 2 * N; j < M; j++) if (q[k * M + j] != 0) return -2;  // quotient degree must be < 2N
    // chunk c, plane k  ->  committed column 2*k + c
    for (int c = 0; c < 2; c++) memcpy(&qcoef[(2 * k + c) * N], &q[k * M + c * N], N * 4);
  }
#pragma omp parallel for
  for (int k = 0; k < 8; k++) coset_eval(&qcoef[k * N], N, &qlde[k * M], lg, shift);
  std::vector<u32> qtree = commit_matrix(qlde.data(), M, 8, p.log_blowup);
  ch.observe_n(merkle_root(qtree.data(), LEAVES), 8);
  memcpy(quot_root_slot, merkle_root(qtree.data(), LEAVES), 32);

  // ---- 3. out-of-domain openings (Horner on coefficients)
  E4 zeta = ch.sample_ext();
  E4 gzeta = e4_mulb(zeta, root_of_unity(log_n));
  if (dump) memcpy(dump->zeta, zeta.c, 16);
  std::vector<E4> ot(WA), otg(WA), oq(8);   // main columns then aux columns
  {
    std::vector<E4> zp(N), gzp(N);  // powers of zeta and g*zeta
    zp[0] = gzp[0] = e4_from(1);
    for (size_t j = 1; j < N; j++) { zp[j] = e4_mul(zp[j - 1], zeta); gzp[j] = e4_mul(gzp[j - 1], gzeta); }
    auto eval = [&](const u32* c, const std::vector<E4>& pw) { E4 acc = e4_zero(); for (size_t j = 0; j < N; j++) acc = e4_add(acc, e4_mulb(pw[j], c[j])); return acc; };
#pragma omp parallel for
    for (size_t k = 0; k < WA; k++) { ot[k] = eval(&coef[k * N], zp); otg[k] = eval(&coef[k * N], gzp); }
    for (int k = 0; k < 8; k++) oq[k] = eval(&qcoef[k * N], zp);
  }
  for (size_t k = 0; k < WA; k++) { memcpy(out, ot[k].c, 16); out += 4; }
  for (size_t k = 0; k < WA; k++) { memcpy(out, otg[k].c, 16); out += 4; }
  for (int k = 0; k < 8; k++) { memcpy(out, oq[k].c, 16); out += 4; }
  {
    u32 od[8];
    hash_tree(out - (2 * WA + 8) * 4, (2 * WA + 8) * 4, od);   // the transcript absorbs the tree hash of the opened values
    ch.observe_n(od, 8);
  }

  // ---- 4. FRI input: batched DEEP quotients on the coset
  E4 af = ch.sample_ext();
  if (dump) memcpy(dump->alpha_fri, af.c, 16);
  std::vector<E4> afp(2 * WA + 8);
  afp[0] = e4_from(1);
  for (size_t k = 1; k < 2 * WA + 8; k++) afp[k] = e4_mul(afp[k - 1], af);
  std::vector<E4> f(M);
  {
    u32 w = root_of_unity(lg);
    std::vector<u32> xs(M);
    u32 x = shift;
    for (size_t i = 0; i < M; i++) { xs[i] = x; x = fmul(x, w); }
    // F(x) = (Rt(x)-A1)/(x-zeta) + alpha^W' (Rt(x)-A2)/(x-g zeta) + alpha^2W' (Rq(x)-A3)/(x-zeta), W' = main + aux columns,
    // Rt(x) = sum_k alpha^k t_k(x), A1 = sum_k alpha^k t_k(zeta), ... (same sum as the per-column DEEP quotients)
    E4 A1 = e4_zero(), A2 = e4_zero(), A3 = e4_zero();
    for (size_t k = 0; k < WA; k++) { A1 = e4_add(A1, e4_mul(afp[k], ot[k])); A2 = e4_add(A2, e4_mul(afp[k], otg[k])); }
    for (int k = 0; k < 8; k++) A3 = e4_add(A3, e4_mul(afp[k], oq[k]));
    std::vector<E4> iz(M), igz(M);
    for (size_t i = 0; i < M; i++) { iz[i] = e4_sub(e4_from(xs[i]), zeta); igz[i] = e4_sub(e4_from(xs[i]), gzeta); }
    const size_t CH = 4096;
#pragma omp parallel for
    for (size_t c0 = 0; c0 < M; c0 += CH) { size_t len = std::min(CH, M - c0); e4_batch_inv(&iz[c0], len); e4_batch_inv(&igz[c0], len); }
#pragma omp parallel for
    for (size_t i = 0; i < M; i++) {
      E4 rt = e4_zero(), rq = e4_zero();
      for (size_t k = 0; k < WA; k++) rt = e4_add(rt, e4_mulb(afp[k], lde[k * M + i]));
      for (int k = 0; k < 8; k++) rq = e4_add(rq, e4_mulb(afp[k], qlde[k * M + i]));
      E4 acc = e4_mul(e4_sub(rt, A1), iz[i]);
      acc = e4_add(acc, e4_mul(afp[WA], e4_mul(e4_sub(rt, A2), igz[i])));
      acc = e4_add(acc, e4_mul(afp[2 * WA], e4_mul(e4_sub(rq, A3), iz[i])));
      f[i] = acc;
    }
  }
  if (dump && dump->fri_input) memcpy(dump->fri_input, f.data(), M * 16);

  // ---- 5. FRI commit phase: a round commits the layer with leaves of `arity` values (the fibre of one point of the layer
  // after the round), samples beta and folds by 2 with beta, beta^2, beta^4 ... (log2(arity) half-folds)
  const size_t R = fri_rounds(log_n);
  std::vector<std::vector<E4>> layers;
  std::vector<std::vector<u32>> trees;
  u32 half_inv = finv(2);
  u32 lshift = shift;  // coset shift of the current layer
  int ll = lg;         // log2 of the current layer length
  for (size_t t = 0; t < R; t++) {
    const u32 la = fri_log_arity(log_n, t), arity = 1u << la;
    size_t n = (size_t)1 << ll, q = n >> la;
    std::vector<u32> tree((2 * q - 1) * 8);
#pragma omp parallel for if (q > 256)
    for (size_t i = 0; i < q; i++) {
      u32 leaf[32];
      for (u32 k = 0; k < arity; k++) memcpy(leaf + 4 * k, f[i + k * q].c, 16);
      hash_elems(leaf, 4 * arity, &tree[i * 8]);
    }
    merkle_build(tree.data(), q);
    const u32* root = merkle_root(tree.data(), q);
    memcpy(out, root, 32); out += 8;
    ch.observe_n(root, 8);
    E4 beta = ch.sample_ext();
    if (dump && dump->betas) memcpy(dump->betas + 4 * t, beta.c, 16);
    layers.push_back(f); trees.push_back(tree);
    for (u32 step = 0; step < la; step++) {
      const size_t h = ((size_t)1 << ll) / 2;
      std::vector<E4> g(h);
      u32 w = root_of_unity(ll), x = lshift;
      std::vector<u32> xs(h);
      for (size_t i = 0; i < h; i++) { xs[i] = x; x = fmul(x, w); }
#pragma omp parallel for if (h > 256)
      for (size_t i = 0; i < h; i++) {
        E4 s = e4_mulb(e4_add(f[i], f[i + h]), half_inv);
        E4 d = e4_mulb(e4_sub(f[i], f[i + h]), finv(fmul(2, xs[i])));
        g[i] = e4_add(s, e4_mul(beta, d));
      }
      f.swap(g);
      lshift = fmul(lshift, lshift);
      beta = e4_mul(beta, beta);
      ll--;
    }
  }
  for (size_t i = 1; i < f.size(); i++) if (!e4_eq(f[i], f[0])) return -3;  // final layer must be constant
  memcpy(out, f[0].c, 16); out += 4;
  ch.observe_n(f[0].c, 4);

  // ---- 6. proof of work
  u32 witness = 0;
  for (;; witness++) {
    Chal c2 = ch;
    c2.observe(witness);
    if (c2.sample_bits(p.pow_bits) == 0) break;
    if (witness == P - 1) return -4;
  }
  ch.observe(witness);
  ch.sample_bits(p.pow_bits);
  *out++ = witness;

  // ---- 7. queries
  for (u32 qi = 0; qi < p.num_queries; qi++) {
    size_t idx = ch.sample_bits(lg);
    const size_t leaf = idx / LR, depth = lg - (p.log_blowup + 1);   // the whole leaf is opened: its 2*B rows, natural order
    for (size_t r = 0; r < LR; r++) for (size_t k = 0; k < W; k++) *out++ = lde[k * M + leaf * LR + r];
    merkle_path(ttree.data(), LEAVES, leaf, out); out += depth * 8;
    for (size_t r = 0; r < LR; r++) for (size_t k = 0; k < AW; k++) *out++ = lde[(W + k) * M + leaf * LR + r];
    merkle_path(atree.data(), LEAVES, leaf, out); out += depth * 8;
    for (size_t r = 0; r < LR; r++) for (int k = 0; k < 8; k++) *out++ = qlde[k * M + leaf * LR + r];
    merkle_path(qtree.data(), LEAVES, leaf, out); out += depth * 8;
    size_t i = idx;
    int ql = lg;
    for (size_t t = 0; t < R; t++) {
      const u32 la = fri_log_arity(log_n, t), arity = 1u << la;
      size_t q = ((size_t)1 << ql) >> la;
      i = i % q;
      for (u32 k = 0; k < arity; k++) { memcpy(out, layers[t][i + k * q].c, 16); out += 4; }
      merkle_path(trees[t].data(), q, i, out); out += (ql - la) * 8;
      ql -= la;
    }
  }
  if ((size_t)(out - proof) != proof_words(p, log_n)) return -5;
  return 0;
}

// ------------------------------------------------------------------ C entry points (ctypes)
extern "C" {
// OpenMP team size of the oracle (launchers such as torchrun export OMP_NUM_THREADS=1); returns the value in effect
int oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}
void oracle_ntt(u32* a, int logn, int inverse) { ntt_inplace(a, logn, inverse != 0); }
void oracle_ntt_batch(u32* a, int ncols, int logn, int inverse) {
#pragma omp parallel for
  for (int k = 0; k < ncols; k++) ntt_inplace(a + ((size_t)k << logn), logn, inverse != 0);
}
// out[ncols][N<<logb] = LDE of in[ncols][N] onto the coset 31*H
void oracle_lde_batch(const u32* in, u32* out, int ncols, int logn, int logb) {
  size_t N = (size_t)1 << logn, M = N << logb;
#pragma omp parallel for
  for (int k = 0; k < ncols; k++) {
    std::vector<u32> c(in + k * N, in + (k + 1) * N);
    ntt_inplace(c.data(), logn, true);
    coset_eval(c.data(), N, out + k * M, logn + logb, ZKIR_BB_GEN);
  }
}
void oracle_poseidon2(u32* states, u64 n) {
#pragma omp parallel for
  for (u64 i = 0; i < n; i++) poseidon2(states + 16 * i);
}
// column-major matrix [ncols][nrows] -> tree ((2*nrows-1)*8 words), returns root
void oracle_merkle_commit(const u32* mat, int ncols, int log_rows, u32* tree, u32* root) {
  size_t n = (size_t)1 << log_rows;
#pragma omp parallel for
  for (size_t i = 0; i < n; i++) {
    std::vector<u32> row(ncols);
    for (int k = 0; k < ncols; k++) row[k] = mat[(size_t)k * n + i];
    hash_elems(row.data(), ncols, tree + 8 * i);
  }
  merkle_build(tree, n);
  memcpy(root, merkle_root(tree, n), 32);
}
// lde: [W + 16 aux][M] (main then aux columns), pub: [4][M]; lookup = {z[4], theta[4]}
void oracle_quotient(const u32* params5, u32 log_n, const u32* lde, const u32* publde, const u32* pv, const u32* lookup8, const u32* io, u64 n_io,
                     const u32* alpha, u32* out) {
  Params p = {params5[0], params5[1], params5[2], params5[3], params5[4]};
  E4 a, z, th; memcpy(a.c, alpha, 16); memcpy(z.c, lookup8, 16); memcpy(th.c, lookup8 + 4, 16);
  const size_t M = (size_t)1 << (log_n + p.log_blowup);
  LookupChallenges lc = make_challenges(z, th);
  lc.sio = io_sum(lc, io, (size_t)n_io);
  quotient_evals(p, log_n, lde, lde + (size_t)p.width * M, publde, pv, lc, a, out);
}
void oracle_public_columns(u32 log_n, const u32* code, u64 n_code, u32* pub) { build_public_columns(log_n, code, (size_t)n_code, pub); }
// aux columns [16][N] for given lookup challenges; returns 1 if the lookups balance
int oracle_aux_columns(u32 log_n, const u32* trace, const u32* code, u64 n_code, const u32* io, u64 n_io, const u32* lookup8, u32* aux) {
  const size_t N = (size_t)1 << log_n;
  std::vector<u32> pub(PW * N);
  build_public_columns(log_n, code, (size_t)n_code, pub.data());
  E4 z, th; memcpy(z.c, lookup8, 16); memcpy(th.c, lookup8 + 4, 16);
  LookupChallenges lc = make_challenges(z, th);
  lc.sio = io_sum(lc, io, (size_t)n_io);
  return build_aux(log_n, trace, pub.data(), lc, aux) ? 1 : 0;
}
void oracle_program_digest(const u32* code, u64 n_code, u32* digest8) { program_digest(code, (size_t)n_code, digest8); }
// one FRI fold: in[n][4] on coset shift*H_n -> out[n/2][4]
void oracle_fri_fold(const u32* in, u32* out, int logn, u32 shift, const u32* beta4) {
  size_t n = (size_t)1 << logn, h = n / 2;
  E4 beta; memcpy(beta.c, beta4, 16);
  u32 w = root_of_unity(logn), x = shift, half_inv = finv(2);
  for (size_t i = 0; i < h; i++) {
    E4 a, b; memcpy(a.c, in + 4 * i, 16); memcpy(b.c, in + 4 * (i + h), 16);
    E4 s = e4_mulb(e4_add(a, b), half_inv);
    E4 d = e4_mulb(e4_sub(a, b), finv(fmul(2, x)));
    E4 g = e4_add(s, e4_mul(beta, d));
    memcpy(out + 4 * i, g.c, 16);
    x = fmul(x, w);
  }
}
// checks every AIR constraint on the (unextended) trace rows, with the aux columns built for the given lookup challenges;
// returns -1 if all hold, -2 if the lookups do not balance, else the index of the first failing constraint (row in *bad_row)
int oracle_check_trace(const u32* trace, u32 log_n, const u32* pv, const u32* code, u64 n_code, const u32* io, u64 n_io, const u32* lookup8, u64* bad_row) {
  size_t N = (size_t)1 << log_n;
  if (log_n < ZKIR_AIR_RANGE_BITS || n_code > N) return -3;
  std::vector<u32> pub(PW * N), aux(AW * N);
  build_public_columns(log_n, code, (size_t)n_code, pub.data());
  E4 z, th; memcpy(z.c, lookup8, 16); memcpy(th.c, lookup8 + 4, 16);
  LookupChallenges lc = make_challenges(z, th);
  lc.sio = io_sum(lc, io, (size_t)n_io);
  const bool balanced = build_aux(log_n, trace, pub.data(), lc, aux.data());
  for (size_t i = 0; i < N; i++) {
    AirRowCtx c;
    c.data = trace; c.aux = aux.data(); c.pub = pub.data(); c.stride = N; c.row = i; c.nxt = (i + 1) % N; c.pv = pv; c.lc = &lc;
    c.is_first = Fp(i == 0); c.is_last = Fp(i == N - 1); c.is_trans = Fp(i != N - 1);
    zkir_air_eval(c);
    for (int k = 0; k < ZKIR_AIR_NUM_CONSTRAINTS; k++) if (!e4_eq(c.vals[k], e4_zero())) { if (bad_row) *bad_row = i; return k; }
  }
  return balanced ? -1 : -2;
}
void oracle_hash_tree(const u32* words, u64 n, u32* digest8) { hash_tree(words, (size_t)n, digest8); }
u64 oracle_proof_words(const u32* params5, u32 log_n) {
  Params p = {params5[0], params5[1], params5[2], params5[3], params5[4]};
  return proof_words(p, log_n);
}
// params5 = {log_blowup, num_queries, pow_bits, width, num_public}; trace column-major [width][1<<log_n]
int oracle_prove(const u32* params5, const u32* trace, u32 log_n, const u32* pv, const u32* code, u64 n_code, const u32* io, u64 n_io, u32* proof) {
  Params p = {params5[0], params5[1], params5[2], params5[3], params5[4]};
  return prove(p, trace, log_n, pv, code, (size_t)n_code, io, (size_t)n_io, proof, nullptr);
}
// challenges20 = alpha, zeta, gamma, lookup z, lookup theta
int oracle_prove_dump(const u32* params5, const u32* trace, u32 log_n, const u32* pv, const u32* code, u64 n_code, const u32* io, u64 n_io, u32* proof, u32* challenges20,
                      u32* lde, u32* aux, u32* quotient, u32* fri_input, u32* betas) {
  Params p = {params5[0], params5[1], params5[2], params5[3], params5[4]};
  Dump d; memset(&d, 0, sizeof(d));
  d.lde = lde; d.aux = aux; d.quotient = quotient; d.fri_input = fri_input; d.betas = betas;
  int rc = prove(p, trace, log_n, pv, code, (size_t)n_code, io, (size_t)n_io, proof, &d);
  if (challenges20) { memcpy(challenges20, d.alpha, 16); memcpy(challenges20 + 4, d.zeta, 16); memcpy(challenges20 + 8, d.alpha_fri, 16);
                      memcpy(challenges20 + 12, d.lookup_z, 16); memcpy(challenges20 + 16, d.lookup_theta, 16); }
  return rc;
}
void oracle_ext_mul(const u32* a, const u32* b, u32* out) { E4 x, y; memcpy(x.c, a, 16); memcpy(y.c, b, 16); E4 r = e4_mul(x, y); memcpy(out, r.c, 16); }
void oracle_ext_inv(const u32* a, u32* out) { E4 x; memcpy(x.c, a, 16); E4 r = e4_inv(x); memcpy(out, r.c, 16); }
void oracle_hash(const u32* in, u64 n, u32* digest) { hash_elems(in, n, digest); }
void oracle_compress(const u32* l, const u32* r, u32* out) { compress(l, r, out); }
}
