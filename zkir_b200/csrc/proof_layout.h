// Word offsets of a protocol v6 proof (docs/PROVER_SPEC.md section 5); shared by the CUDA prover and the host verifier.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "zkir_b200.h"
#include "air_profiles_generated.h"

namespace zkir {

static const uint32_t PROOF_MAGIC = 0x5A4B5052u, PROOF_VERSION = 6u;
static const uint32_t QW = 8;                      // quotient columns: 4 ext planes x 2 chunks, column = 2*plane + chunk
// The AIR profile of a proof is its trace width (docs/PROVER_SPEC.md section 3.7): aux (LogUp) columns committed after the lookup
// challenges, public columns (never committed), constraints
inline bool profile_is_full(uint32_t width) { return width == ZKIR_PROFILE_FULL_WIDTH; }
inline bool profile_known(uint32_t width) { return width == ZKIR_PROFILE_CORE_WIDTH || width == ZKIR_PROFILE_FULL_WIDTH; }
inline uint32_t profile_aux_width(uint32_t width) { return profile_is_full(width) ? ZKIR_PROFILE_FULL_AUX : ZKIR_PROFILE_CORE_AUX; }
inline uint32_t profile_pub_width(uint32_t width) { return profile_is_full(width) ? ZKIR_PROFILE_FULL_PUB : ZKIR_PROFILE_CORE_PUB; }
inline uint32_t profile_constraints(uint32_t width) { return profile_is_full(width) ? ZKIR_PROFILE_FULL_CONSTRAINTS : ZKIR_PROFILE_CORE_CONSTRAINTS; }
static const uint32_t ZKIR_NUM_PUBLIC_VALUES = 5;   // entry_pc, num_cycles, exit_lo, exit_hi, halted (both profiles)
static const uint32_t ZKIR_RANGE_BITS = 10;         // the 1024-row range table sets the minimum trace length (both profiles)

// Merkle leaves of the three LDE matrices hold 2^log_leaf_rows = 2*B consecutive natural-order rows (two adjacent trace points on
// all B cosets): a quarter of the tree compressions at B = 2 for three more opened rows per query (docs/PROVER_SPEC.md section 4.1)
inline uint32_t log_leaf_rows(uint32_t log_blowup) { return log_blowup + 1; }

struct Layout {  // proof word offsets
  uint32_t log_n, log_m, width, aw, wa, np, nq, R, log_lr, depth;   // depth = levels of a matrix tree = log_m - log_lr
  size_t pv, troot, aroot, qroot, open_t, open_tg, open_q, fri_roots, final_, pow_, queries, per_query, total;
  // inside one query: trace row, its path, aux row, its path, quotient row, its path, then the FRI rounds
  size_t q_trow, q_tpath, q_arow, q_apath, q_qrow, q_qpath, q_fri;
};
inline Layout make_layout(const zkir_params* p, uint32_t log_n) {
  Layout L;
  L.log_n = log_n; L.log_m = log_n + p->log_blowup; L.width = p->width; L.aw = profile_aux_width(p->width); L.wa = p->width + L.aw; L.np = p->num_public; L.nq = p->num_queries;
  L.log_lr = log_leaf_rows(p->log_blowup); L.depth = L.log_m - L.log_lr;
  L.R = log_n / 3 + (log_n % 3 ? 1 : 0);   // FRI rounds: log_n / 3 that fold by 8, one more by 2^(log_n mod 3) (docs/PROVER_SPEC.md 4.6)
  size_t o = 8;
  L.pv = o; o += L.np;
  L.troot = o; o += 8;
  L.aroot = o; o += 8;
  L.qroot = o; o += 8;
  L.open_t = o; o += 4 * (size_t)L.wa;
  L.open_tg = o; o += 4 * (size_t)L.wa;
  L.open_q = o; o += 4 * QW;
  L.fri_roots = o; o += 8 * (size_t)L.R;
  L.final_ = o; o += 4;
  L.pow_ = o; o += 1;
  L.queries = o;
  size_t q = 0;
  const size_t lr = (size_t)1 << L.log_lr;
  L.q_trow = q; q += lr * L.width;             // the 2*B rows of the opened leaf, natural order
  L.q_tpath = q; q += 8 * (size_t)L.depth;
  L.q_arow = q; q += lr * L.aw;
  L.q_apath = q; q += 8 * (size_t)L.depth;
  L.q_qrow = q; q += lr * QW;
  L.q_qpath = q; q += 8 * (size_t)L.depth;
  L.q_fri = q;
  for (uint32_t t = 0, ll = L.log_m; t < L.R; t++) { const uint32_t la = t < log_n / 3 ? 3 : log_n % 3; q += (4u << la) + 8 * (size_t)(ll - la); ll -= la; }
  L.per_query = q;
  L.total = o + q * L.nq;
  return L;
}

}  // namespace zkir
