// CPU verifier for zkir-b200 proofs (product host code; shares no source with oracle/).
//
// The reference has no proof type and no verifier (zkir-runtime/src/lib.rs:29-62); this implements the verifier
// side of docs/PROVER_SPEC.md: replay the Fiat-Shamir transcript, check the AIR identity at zeta
// (air_generated.h instantiated over ext4), recompute the DEEP/FRI input at every query from the opened rows,
// check all Merkle paths and the FRI folding chain down to the constant.
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "zkir_b200.h"
#include "../constants_generated.h"
#include "../air_profile.h"   // this file is compiled once per AIR profile; zkir_b200_verify (core build) dispatches on the width
#include "../air_profiles_generated.h"
#include <algorithm>

std::string& zkir_host_error();
#define g_verify_error zkir_host_error()
namespace {
typedef uint32_t u32;
typedef uint64_t u64;
const u32 P = ZKIR_BB_P;

inline u32 add(u32 a, u32 b) { u64 s = (u64)a + b; return (u32)(s >= P ? s - P : s); }
inline u32 sub(u32 a, u32 b) { return a >= b ? a - b : (u32)((u64)a + P - b); }
inline u32 mul(u32 a, u32 b) { return (u32)((u64)a * b % P); }
u32 pw(u32 a, u64 e) { u32 r = 1; for (; e; e >>= 1, a = mul(a, a)) if (e & 1) r = mul(r, a); return r; }
inline u32 inv(u32 a) { return pw(a, P - 2); }

struct X4 {  // F_p[X]/(X^4-11)
  u32 c[4];
  X4() { c[0] = c[1] = c[2] = c[3] = 0; }
  explicit X4(u32 b) { c[0] = b; c[1] = c[2] = c[3] = 0; }
};
inline X4 operator+(const X4& a, const X4& b) { X4 r; for (int i = 0; i < 4; i++) r.c[i] = add(a.c[i], b.c[i]); return r; }
inline X4 operator-(const X4& a, const X4& b) { X4 r; for (int i = 0; i < 4; i++) r.c[i] = sub(a.c[i], b.c[i]); return r; }
inline X4 operator*(const X4& a, const X4& b) {
  u64 t[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t[i + j] = (t[i + j] + (u64)a.c[i] * b.c[j]) % P;
  X4 r;
  for (int i = 0; i < 4; i++) r.c[i] = (u32)t[i];
  for (int i = 4; i < 7; i++) r.c[i - 4] = (u32)((r.c[i - 4] + (u64)ZKIR_EXT_W * t[i]) % P);
  return r;
}
inline X4 scale(const X4& a, u32 b) { X4 r; for (int i = 0; i < 4; i++) r.c[i] = mul(a.c[i], b); return r; }
inline bool eq(const X4& a, const X4& b) { return !memcmp(a.c, b.c, 16); }
X4 xpow(X4 a, u64 e) { X4 r(1); for (; e; e >>= 1, a = a * a) if (e & 1) r = r * a; return r; }
// inverse by solving the 4x4 linear system  M_a * y = e_0  (Gaussian elimination over F_p)
bool xinv(const X4& a, X4* out) {
  u32 m[4][5];
  for (int col = 0; col < 4; col++) {  // column `col` of M_a is a * X^col
    X4 b; b.c[col] = 1;
    X4 pr = a * b;
    for (int row = 0; row < 4; row++) m[row][col] = pr.c[row];
  }
  for (int row = 0; row < 4; row++) m[row][4] = row == 0;
  for (int col = 0; col < 4; col++) {
    int piv = -1;
    for (int r = col; r < 4; r++) if (m[r][col]) { piv = r; break; }
    if (piv < 0) return false;
    if (piv != col) for (int k = 0; k < 5; k++) std::swap(m[piv][k], m[col][k]);
    u32 iv = inv(m[col][col]);
    for (int k = 0; k < 5; k++) m[col][k] = mul(m[col][k], iv);
    for (int r = 0; r < 4; r++) if (r != col && m[r][col]) {
      u32 f = m[r][col];
      for (int k = 0; k < 5; k++) m[r][k] = sub(m[r][k], mul(f, m[col][k]));
    }
  }
  for (int i = 0; i < 4; i++) out->c[i] = m[i][4];
  return true;
}

// ---- Poseidon2 width 16
inline u32 sbox(u32 x) { u32 x2 = mul(x, x), x4 = mul(x2, x2); return mul(mul(x4, x2), x); }
void ext_layer(u32* s) {
  u32 t[16];
  for (int g = 0; g < 4; g++) for (int j = 0; j < 4; j++) {
    const u32* x = s + 4 * g;
    u64 v = 2ull * x[j] + 3ull * x[(j + 1) & 3] + x[(j + 2) & 3] + x[(j + 3) & 3];
    t[4 * g + j] = (u32)(v % P);
  }
  for (int i = 0; i < 16; i++) {
    u64 v = (u64)t[i] + t[i & 3] + t[4 + (i & 3)] + t[8 + (i & 3)] + t[12 + (i & 3)];
    s[i] = (u32)(v % P);
  }
}
void permute(u32* s) {
  ext_layer(s);
  for (int r = 0; r < 8; r++) {
    if (r == 4) {
      for (int k = 0; k < ZKIR_P2_RP; k++) {
        s[0] = sbox(add(s[0], ZKIR_P2_RC_INT[k]));
        u64 sum = 0;
        for (int i = 0; i < 16; i++) sum += s[i];
        u32 sm = (u32)(sum % P);
        for (int i = 0; i < 16; i++) s[i] = add(mul(s[i], ZKIR_P2_DIAG[i]), sm);
      }
    }
    for (int i = 0; i < 16; i++) s[i] = sbox(add(s[i], ZKIR_P2_RC_EXT[16 * r + i]));
    ext_layer(s);
  }
}
void hash_n(const u32* in, size_t n, u32* d) {
  u32 s[16] = {0};
  for (size_t i = 0; i < n; i += 8) { for (size_t k = 0; k < 8 && i + k < n; k++) s[k] = in[i + k]; permute(s); }
  memcpy(d, s, 32);
}
void compress2(const u32* l, const u32* r, u32* d) { u32 s[16]; memcpy(s, l, 32); memcpy(s + 8, r, 32); permute(s); memcpy(d, s, 32); }
// transcript digest of a long message (docs/PROVER_SPEC.md section 2, `hash_tree`): 8-word chunks are leaves, zero digests pad
// the leaf level to a power of two, the Merkle root is the digest
void hash_tree(const u32* words, size_t n, u32* digest) {
  size_t chunks = (n + 7) / 8, leaves = 1;
  while (leaves < chunks) leaves <<= 1;
  std::vector<u32> lvl(leaves * 8, 0);
  for (size_t c = 0; c < chunks; c++) hash_n(words + 8 * c, n - 8 * c < 8 ? n - 8 * c : 8, &lvl[8 * c]);
  for (size_t m = leaves; m > 1; m >>= 1)
    for (size_t i = 0; i < m / 2; i++) { u32 d[8]; compress2(&lvl[16 * i], &lvl[16 * i + 8], d); memcpy(&lvl[8 * i], d, 32); }
  memcpy(digest, lvl.data(), 32);
}
bool check_path(const u32* leaf_digest, u64 idx, const u32* path, u32 depth, const u32* root) {
  u32 cur[8]; memcpy(cur, leaf_digest, 32);
  for (u32 l = 0; l < depth; l++, idx >>= 1) {
    u32 nxt[8];
    if (idx & 1) compress2(path + 8 * l, cur, nxt); else compress2(cur, path + 8 * l, nxt);
    memcpy(cur, nxt, 32);
  }
  return !memcmp(cur, root, 32);
}

struct Challenger {
  u32 st[16], in[8], out[8]; int nin, nout;
  Challenger() { memset(this, 0, sizeof(*this)); }
  void duplex() { for (int i = 0; i < nin; i++) st[i] = in[i]; nin = 0; permute(st); memcpy(out, st, 32); nout = 8; }
  void observe(u32 x) { nout = 0; in[nin++] = x; if (nin == 8) duplex(); }
  void observe(const u32* x, size_t n) { for (size_t i = 0; i < n; i++) observe(x[i]); }
  u32 sample() { if (nin || !nout) duplex(); return out[--nout]; }
  X4 sample_ext() { X4 r; for (int i = 0; i < 4; i++) r.c[i] = sample(); return r; }
  u32 bits(u32 b) { u32 v = sample(); return b >= 32 ? v : v & ((1u << b) - 1); }
};

struct AirAtZeta {  // air_generated.h context over ext4: base and ext values are both X4 here
  typedef X4 F; typedef X4 X;
  const X4 *loc, *nxt;        // opened main columns [0, W) then aux columns [W, W + A) at zeta / g*zeta
  const X4* pub;              // public columns at zeta, computed by the verifier
  const u32* pv;
  X4 zc, thp[ZKIR_AIR_NUM_THETA + 1], sio_v;       // lookup challenges z, theta^k; the public I/O transcript's sum
  X4 is_first, is_last, is_trans, alpha, acc;
  X4 L(int i) const { return loc[i]; }
  X4 N(int i) const { return nxt[i]; }
  X4 A(int i) const { return loc[ZKIR_AIR_WIDTH + i]; }
  X4 AN(int i) const { return nxt[ZKIR_AIR_WIDTH + i]; }
  X4 P(int i) const { return pub[i]; }
  X4 PV(int i) const { return X4(pv[i]); }
  X4 K(u32 k) const { return X4(k); }
  X4 z() const { return zc; }
  X4 th(int k) const { return thp[k]; }
  X4 xf(const X4& a) const { return a; }
  X4 sio() const { return sio_v; }
  X4 x4(const X4& a, const X4& b, const X4& c, const X4& d) const {   // a + X b + X^2 c + X^3 d: the ext value of four base polynomials
    X4 e1, e2, e3; e1.c[1] = 1; e2.c[2] = 1; e3.c[3] = 1;
    return a + e1 * b + e2 * c + e3 * d;
  }
  void fence() const {}
  void emit(int, const X4& v) { acc = acc * alpha + v; }  // Horner: sum_i alpha^(K-1-i) c_i
  void emit_x(int, const X4& v) { acc = acc * alpha + v; }
};

// ROM decode of one instruction word, as the AIR's ROM lookup sees it (docs/PROVER_SPEC.md section 3.3): the fields of the word's
// format (zkir-assembler/src/encoder.rs:98-151) packed as opcode | rd << 7 | rs1 << 11 | rs2 << 15, and the signed immediate mod p
inline int32_t sx(u32 v, int bits) { int sh = 32 - bits; return ((int32_t)(v << sh)) >> sh; }
void rom_entry(u32 w, u32* dec, u32* imm) {
  u32 f[5];
  int64_t im = 0; u32 rd = 0, rs1 = 0, rs2 = 0;
  const u32 op = w & 0x7F;
  if (zkir_decode(w, f) != 0) { *dec = op | ((w >> 7) & 15) << 7 | ((w >> 11) & 15) << 11 | ((w >> 15) & 15) << 15; *imm = 0; return; }  // undefined opcode: matches no row
  const bool stype = (op >= 0x38 && op <= 0x3B) || (op >= 0x40 && op <= 0x45);
  const bool rtype = !stype && op != 0x48 && op != 0x50 && op != 0x51 && !(op == 0x08 || (op >= 0x13 && op <= 0x15) || (op >= 0x30 && op <= 0x35) || op == 0x49) &&
                     !(op >= 0x1B && op <= 0x1D);
  if (stype) { rs1 = f[1]; rs2 = f[2]; im = (int32_t)f[4]; }
  else if (rtype) { rd = f[1]; rs1 = f[2]; rs2 = f[3]; }
  else if (op == 0x48) { rd = f[1]; im = (int32_t)f[4]; }
  else if (op == 0x50 || op == 0x51) {}
  else { rd = f[1]; rs1 = f[2]; im = (int32_t)f[4]; }     // I-type and shift-immediate
  *dec = op | rd << 7 | rs1 << 11 | rs2 << 15;
  *imm = im < 0 ? (u32)((int64_t)P + im) : (u32)im;
}
// transcript digest of the program: hash_tree over {n_code, lo16(word_0), hi16(word_0), ...}
void program_digest(const u32* code, size_t n_code, u32* digest) {
  std::vector<u32> h(2 * n_code + 1);
  h[0] = (u32)n_code;
  for (size_t i = 0; i < n_code; i++) { h[1 + 2 * i] = code[i] & 0xFFFF; h[2 + 2 * i] = code[i] >> 16; }
  hash_tree(h.data(), h.size(), digest);
}

bool fail(const char* m) { g_verify_error = m; return false; }

// transcript digest of the public I/O transcript: hash_tree over {n_io, clk_0, kind_0, lo_0, hi_0, ...}
void io_digest(const u32* io, size_t n_io, u32* digest) {
  std::vector<u32> h(4 * n_io + 1);
  h[0] = (u32)n_io;
  for (size_t i = 0; i < 4 * n_io; i++) h[1 + i] = io[i] % P;
  hash_tree(h.data(), h.size(), digest);
}

bool verify(const zkir_params* p, const u32* w, size_t nwords, const u32* pv_in, const u32* code, size_t n_code, const u32* io, size_t n_io) {
  if (!p || !w) return fail("null argument");
  if (p->width != ZKIR_AIR_WIDTH || p->num_public != ZKIR_AIR_NUM_PUBLIC || p->log_blowup < 1) return fail("unsupported params");
  if (nwords < 8) return fail("proof too short");
  if (w[0] != 0x5A4B5052u || w[1] != 6) return fail("bad magic/version");
  if ((!code && n_code) || (!io && n_io)) return fail("null program / I/O transcript");
  const u32 log_n = w[2];
  if (w[3] != p->width || w[4] != p->log_blowup || w[5] != p->num_queries || w[6] != p->pow_bits || w[7] != p->num_public)
    return fail("proof header does not match params");
  if (log_n < ZKIR_AIR_RANGE_BITS || log_n + p->log_blowup > 27) return fail("bad log_n");
  if (n_code > (1ull << log_n)) return fail("program does not fit the trace");
  if (nwords * 4 != zkir_b200_proof_size(p, log_n)) return fail("proof length mismatch");
  const u32 W = p->width, AW = ZKIR_AIR_AUX_WIDTH, WA = W + AW, np = p->num_public, lg = log_n + p->log_blowup, R = log_n / 3 + (log_n % 3 ? 1 : 0), QW = 8;  // R FRI rounds: fold by 8, the last by 2^(log_n mod 3)
  const u64 N = 1ull << log_n, M = 1ull << lg;
  for (size_t i = 8; i < nwords; i++) if (w[i] >= P) return fail("non-canonical field element");
  const u32* q = w + 8;
  const u32* pv = q; q += np;
  if (pv_in) for (u32 i = 0; i < np; i++) if (pv[i] != pv_in[i]) return fail("public values differ");
  // public values that are not bound by constraints alone: halted is a flag, and a run that did not halt has no exit code
  if (pv[4] > 1 || (pv[4] == 0 && (pv[2] || pv[3]))) return fail("inconsistent halt flag / exit code");
  const u32* troot = q; q += 8;
  const u32* aroot = q; q += 8;
  const u32* qroot = q; q += 8;
  const u32* open = q;
  std::vector<X4> ot(WA), otg(WA), oq(QW);   // main columns then aux columns
  for (u32 k = 0; k < WA; k++) { memcpy(ot[k].c, q, 16); q += 4; }
  for (u32 k = 0; k < WA; k++) { memcpy(otg[k].c, q, 16); q += 4; }
  for (u32 k = 0; k < QW; k++) { memcpy(oq[k].c, q, 16); q += 4; }
  const u32* fri_roots = q; q += 8 * R;
  X4 final_v; memcpy(final_v.c, q, 16); const u32* final_w = q; q += 4;
  const u32 witness = *q++;

  // ---- transcript
  Challenger ch;
  const u32 hdr[7] = {log_n, W, AW, p->log_blowup, p->num_queries, p->pow_bits, np};
  u32 pdig[8];
  program_digest(code, n_code, pdig);
  ch.observe(hdr, 7); ch.observe(pv, np); ch.observe(pdig, 8);
  io_digest(io, n_io, pdig);
  ch.observe(pdig, 8); ch.observe(troot, 8);
  const X4 lz = ch.sample_ext(), ltheta = ch.sample_ext();   // lookup challenges, drawn before the aux columns are committed
  ch.observe(aroot, 8);
  const X4 alpha = ch.sample_ext();
  ch.observe(qroot, 8);
  const X4 zeta = ch.sample_ext();
  u32 open_digest[8];
  hash_tree(open, (2 * WA + QW) * 4, open_digest);
  ch.observe(open_digest, 8);
  const X4 afri = ch.sample_ext();
  std::vector<X4> betas(R);
  for (u32 r = 0; r < R; r++) { ch.observe(fri_roots + 8 * r, 8); betas[r] = ch.sample_ext(); }
  ch.observe(final_w, 4);
  ch.observe(witness);
  if (ch.bits(p->pow_bits) != 0) return fail("proof-of-work check failed");

  // ---- AIR identity at zeta:  C(zeta) = Z_H(zeta) * (q0(zeta) + zeta^N q1(zeta))
  const u32 g = ZKIR_BB_ROOTS[log_n], g_inv = inv(g);
  const X4 zN = xpow(zeta, N), zh = zN - X4(1);
  X4 i1, i2;
  if (!xinv(zeta - X4(1), &i1) || !xinv(zeta - X4(g_inv), &i2)) return fail("zeta hits the trace domain");
  // public columns at zeta by the barycentric formula over H_N:  P(zeta) = (zeta^N - 1)/N * sum_i v_i w^i / (zeta - w^i);
  // only the first max(1024, n_code) rows are non-zero, except the ROM's decoded-word column which is 127 on every unused row:
  // it is evaluated as the constant 127 plus the interpolant of (dec_i - 127) over the program rows
  // a column is its default value (the rows past the tables and the program: zkir_public_row(~0)) plus the interpolant of the
  // differences over the first rows
  X4 pubz[ZKIR_AIR_PUB_WIDTH];
  {
    const size_t rows = std::min<size_t>((size_t)zkir_public_rows(p->width, n_code), (size_t)N);
    const X4 scale_all = scale(zh, inv((u32)(N % P)));
    u32 dflt[ZKIR_AIR_PUB_WIDTH], row[ZKIR_AIR_PUB_WIDTH];
    zkir_public_row(p->width, ~0ull, code, n_code, dflt);
    u32 wi = 1;
    for (size_t i = 0; i < rows; i++) {
      X4 d;
      if (!xinv(zeta - X4(wi), &d)) return fail("zeta hits the trace domain");
      const X4 li = scale(d, wi);    // w^i / (zeta - w^i)
      zkir_public_row(p->width, i, code, n_code, row);
      for (int k = 0; k < ZKIR_AIR_PUB_WIDTH; k++) if (row[k] != dflt[k]) pubz[k] = pubz[k] + scale(li, sub(row[k], dflt[k]));
      wi = mul(wi, g);
    }
    for (int k = 0; k < ZKIR_AIR_PUB_WIDTH; k++) pubz[k] = pubz[k] * scale_all + X4(dflt[k]);
  }
  AirAtZeta c;
  c.loc = ot.data(); c.nxt = otg.data(); c.pub = pubz; c.pv = pv; c.alpha = alpha;
  c.zc = lz; c.thp[0] = X4(1); c.thp[1] = ltheta;
  for (int k = 2; k <= ZKIR_AIR_NUM_THETA; k++) c.thp[k] = c.thp[k - 1] * ltheta;
  // the table side of the I/O bus is public: S_io = sum_e 1 / (z - (3 + theta clk + theta^2 kind + theta^3 lo + theta^4 hi))
  for (size_t e = 0; e < n_io; e++) {
    X4 fp(3), d;
    for (int k = 0; k < 4; k++) fp = fp + scale(c.thp[k + 1], io[4 * e + k] % P);
    if (!xinv(lz - fp, &d)) return fail("lookup challenge hits an I/O event");
    c.sio_v = c.sio_v + d;
  }
  c.is_first = zh * i1; c.is_last = zh * i2; c.is_trans = zeta - X4(g_inv);
  zkir_air_eval(c);
  X4 xp[4]; for (int k = 0; k < 4; k++) { xp[k] = X4(); xp[k].c[k] = 1; }  // basis 1, X, X^2, X^3
  X4 qz;
  for (int chunk = 0; chunk < 2; chunk++) {
    X4 v;
    for (int k = 0; k < 4; k++) v = v + xp[k] * oq[2 * k + chunk];
    qz = qz + (chunk ? zN * v : v);
  }
  if (!eq(c.acc, zh * qz)) return fail("constraint identity fails at zeta");

  // ---- queries
  std::vector<X4> afp(2 * WA + 1);
  afp[0] = X4(1);
  for (u32 k = 1; k <= 2 * WA; k++) afp[k] = afp[k - 1] * afri;
  X4 A1, A2, A3;
  for (u32 k = 0; k < WA; k++) { A1 = A1 + afp[k] * ot[k]; A2 = A2 + afp[k] * otg[k]; }
  for (u32 k = 0; k < QW; k++) A3 = A3 + afp[k] * oq[k];
  const X4 gzeta = scale(zeta, g);
  const u32 wM = ZKIR_BB_ROOTS[lg], half = inv(2);
  for (u32 qi = 0; qi < p->num_queries; qi++) {
    const u64 idx = ch.bits(lg);
    // a matrix leaf holds LR = 2*B consecutive natural-order rows: all of them are opened, row idx mod LR is the queried one
    const u32 LR = 2u << p->log_blowup, depth = lg - (p->log_blowup + 1);
    const u64 leaf = idx / LR;
    const u32 sub = (u32)(idx % LR);
    const u32* trows = q; q += LR * W;
    const u32* tpath = q; q += 8 * depth;
    const u32* arows = q; q += LR * AW;
    const u32* apath = q; q += 8 * depth;
    const u32* qrows = q; q += LR * QW;
    const u32* qpath = q; q += 8 * depth;
    const u32 *trow = trows + sub * W, *arow = arows + sub * AW, *qrow = qrows + sub * QW;
    u32 d[8];
    hash_n(trows, LR * W, d);
    if (!check_path(d, leaf, tpath, depth, troot)) return fail("trace Merkle path");
    hash_n(arows, LR * AW, d);
    if (!check_path(d, leaf, apath, depth, aroot)) return fail("aux Merkle path");
    hash_n(qrows, LR * QW, d);
    if (!check_path(d, leaf, qpath, depth, qroot)) return fail("quotient Merkle path");
    const u32 x = mul(ZKIR_BB_GEN, pw(wM, idx));
    X4 rt, rq, iz, igz;
    for (u32 k = 0; k < W; k++) rt = rt + scale(afp[k], trow[k]);
    for (u32 k = 0; k < AW; k++) rt = rt + scale(afp[W + k], arow[k]);
    for (u32 k = 0; k < QW; k++) rq = rq + scale(afp[k], qrow[k]);
    if (!xinv(X4(x) - zeta, &iz) || !xinv(X4(x) - gzeta, &igz)) return fail("zeta on the LDE coset");
    X4 v = (rt - A1) * iz + afp[WA] * ((rt - A2) * igz) + afp[2 * WA] * ((rq - A3) * iz);
    u64 i = idx;
    u32 lshift = ZKIR_BB_GEN;
    u32 ll = lg;  // log2 of the layer length
    for (u32 t = 0; t < R; t++) {
      const u32 la = t < log_n / 3 ? 3 : log_n % 3, arity = 1u << la;
      const u64 qn = (1ull << ll) >> la;            // leaves of this layer; the opened values sit at i + k*qn
      const u32 pos = (u32)(i / qn);
      i %= qn;
      X4 a[8];
      for (u32 k = 0; k < arity; k++) memcpy(a[k].c, q + 4 * k, 16);
      const u32* vals = q; q += 4 * arity;
      const u32* path = q; q += 8 * (ll - la);
      if (!eq(a[pos], v)) return fail("FRI layer value does not match the folded value");
      hash_n(vals, 4 * arity, d);
      if (!check_path(d, i, path, ll - la, fri_roots + 8 * t)) return fail("FRI Merkle path");
      // log2(arity) half-folds: step s pairs value k with k + arity/2^(s+1) at the point (x_i * w_arity^k)^(2^s), challenge beta^(2^s)
      u32 pts[4];
      const u32 xi = mul(lshift, pw(ZKIR_BB_ROOTS[ll], i));
      for (u32 k = 0; k < arity / 2; k++) pts[k] = mul(xi, pw(ZKIR_BB_ROOTS[la], k));
      X4 bp = betas[t];
      for (u32 s = 0; s < la; s++) {
        const u32 half_n = arity >> (s + 1);
        for (u32 k = 0; k < half_n; k++) {
          a[k] = scale(a[k] + a[k + half_n], half) + bp * scale(a[k] - a[k + half_n], inv(mul(2, pts[k])));
          pts[k] = mul(pts[k], pts[k]);
        }
        bp = bp * bp;
      }
      v = a[0];
      for (u32 k = 0; k < la; k++) lshift = mul(lshift, lshift);
      ll -= la;
    }
    if (!eq(v, final_v)) return fail("FRI final value mismatch");
  }
  if ((size_t)(q - w) != nwords) return fail("trailing data");
  return true;
}
}  // namespace

extern "C" {
#ifdef ZKIR_PROFILE_FULL
// the full-profile build of this file exports only its verifier; zkir_b200_verify (core build) calls it for full-width proofs
int zkir_verify_words_full(const zkir_params* p, const uint32_t* w, size_t nwords, const uint32_t* public_values, const uint32_t* code, size_t n_code,
                           const uint32_t* io_events, size_t n_io) {
  return verify(p, w, nwords, public_values, code, n_code, io_events, n_io) ? 0 : ZKIR_ERR_VERIFY;
}
#else
int zkir_verify_words_full(const zkir_params*, const uint32_t*, size_t, const uint32_t*, const uint32_t*, size_t, const uint32_t*, size_t);
int zkir_b200_verify(const zkir_params* p, const uint8_t* proof, size_t len, const uint32_t* public_values, const uint32_t* code, size_t n_code,
                     const uint32_t* io_events, size_t n_io) {
  g_verify_error.clear();
  if (!p || !proof || len % 4) { g_verify_error = "bad proof buffer"; return ZKIR_ERR_VERIFY; }
  std::vector<u32> w(len / 4);
  memcpy(w.data(), proof, len);
  if (p->width == ZKIR_PROFILE_FULL_WIDTH) return zkir_verify_words_full(p, w.data(), w.size(), public_values, code, n_code, io_events, n_io);
  return verify(p, w.data(), w.size(), public_values, code, n_code, io_events, n_io) ? 0 : ZKIR_ERR_VERIFY;
}
void zkir_program_digest(const uint32_t* code, size_t n_code, uint32_t digest8[8]) { program_digest(code, n_code, digest8); }
void zkir_io_digest(const uint32_t* io_events, size_t n_io, uint32_t digest8[8]) { io_digest(io_events, n_io, digest8); }
// the width-16 Poseidon2 permutation (canonical values in and out): SYS_POSEIDON2 of the interpreter (vm.cc) uses it
void zkir_host_poseidon2_permute(uint32_t* state16) { permute(state16); }
// ROM entry of one code word as the AIR's ROM lookup sees it (used by the prover to build the public ROM columns)
void zkir_rom_entry(uint32_t word, uint32_t* dec, uint32_t* imm) { rom_entry(word, dec, imm); }
#endif  // !ZKIR_PROFILE_FULL
}
