// Internal launcher interface between the .cu translation units of libzkir_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bb.cuh"

namespace zkir {

struct ChalState {  // duplex challenger, Poseidon2 width 16 / rate 8 (docs/PROVER_SPEC.md "Transcript")
  u32 sponge[16];
  u32 inbuf[8];
  u32 outbuf[8];
  u32 n_in, n_out;
};

// ---- ntt.cu
struct NttTables;
NttTables* ntt_tables_create(cudaStream_t st, u64* launch_counter);
void ntt_tables_destroy(NttTables*);
u32* ntt_powers_table(NttTables*, u32 base_canon, u32 c0_canon, u64 n);  // cached c0*base^i, Montgomery
int ntt_run(NttTables* tb, const u32* d_in, u64 in_col_stride, u32* d_out, u64 out_col_stride, u32* tmp, u64 tmp_words,
            u32 n_cols, int log_n, bool inverse, int log_pad, const u32* in_scale, const u32* out_scale,
            u32 out_const_mont, bool use_out_const, cudaStream_t st);

#define ZKIR_MAX_SHARDS 64
struct PeerPtrs { u32* p[ZKIR_MAX_SHARDS]; };   // base of every rank's matrix as seen from this device (own entry included)

// ---- ntt_fast.cu (register radix-32 tiles, digit-reversed coefficients, coset-major LDE)
struct FastNtt;
struct FastPlan { int log_n, nd, d[4]; };   // digit bits, top to bottom in the memory position
FastNtt* fast_ntt_create(cudaStream_t st, u64* launch_counter);
void fast_ntt_destroy(FastNtt*);
bool fast_plan(int log_n, FastPlan* out);   // false when log_n < 8 (callers use the generic path)
u64 fast_plan_coef_index(const FastPlan& pl, u64 pos);
int fast_intt(FastNtt* f, const FastPlan& pl, const u32* in, u64 in_col, u32* coef, u64 coef_col, u32 n_cols, u32 c0_canon,
              const uint2* coef_tab, u32 split_log, u32 split_max, u32 split_extra, u32* split_out, u64 split_out_col, cudaStream_t st);
// Sharded proofs (peers != nullptr): the LAST pass stores every output row straight into the matrix of the rank that owns the row
// (rows [g*nj, (g+1)*nj) of every coset, plus the halo row (g+1)*nj mod N) through the peer pointers instead of into `out`; `out`
// must lie inside the matrix whose base is peers->p[me].  Requires G <= 2^(top digit), see fast_coset_ntt_can_fuse.
int fast_coset_ntt(FastNtt* f, const FastPlan& pl, const u32* coef, u64 coef_col, u32* out, u64 out_col, u32 n_cols, u32 nz,
                   u32 base_canon, u32 zroot_canon, u32 c0_canon, cudaStream_t st, const PeerPtrs* peers = nullptr, u32 me = 0, u32 G = 1);
bool fast_coset_ntt_can_fuse(const FastPlan& pl, u32 G);
int fast_ntt_natural(FastNtt* f, int log_n, bool inverse, u32 coset_shift, const u32* in, u64 in_col, u32* tmp, u32* out, u64 out_col,
                     u32 n_cols, cudaStream_t st);
int launch_coset_reorder(const u32* in, u32* out, u32 n_cols, u32 log_n, u32 log_b, int to_natural, cudaStream_t st, u64* launches);
const uint2* fast_scale_table(FastNtt* f, const FastPlan& pl, u32 base_canon, u32 c0_canon);  // [N] pairs c0*base^k(pos)

// ---- poseidon2.cu
int poseidon2_init_constants();
int launch_permute(u32* d_states, u64 n, bool canonical_io, cudaStream_t st, u64* launches);
// matrix rows are coset-major (kernels.h: see ntt_fast.cu), leaves natural; log_b = 0 for a natural-order matrix
// first_leaf / seg_leaves != 0: only that range of (natural-order) leaves, for a proof sharded over several GPUs
int launch_leaf_hash(const u32* mat, u64 col_stride, u32 n_cols, u64 n_rows, u32 log_b, u32* digests, cudaStream_t st, u64* launches,
                     u64 first_leaf = 0, u64 seg_leaves = 0);
// the prover's leaves: 2*B consecutive natural-order rows per leaf (two adjacent trace points on all cosets); leaves [first, first + n)
int launch_leaf_hash_rows(const u32* mat, u64 col_stride, u32 n_cols, u32 log_n, u32 log_b, u32* digests, cudaStream_t st, u64* launches,
                          u64 first_leaf, u64 n_leaves);
// optional fused Fiat-Shamir step on the root: copy to root_dst, observe, sample n_sample elements into sample_out
// pair_layer != nullptr: the leaves are FRI leaves hash(f[i] || f[i + n_leaves]) of that ext4 layer and are computed here too
int launch_merkle_levels(u32* tree, u64 n_leaves, cudaStream_t st, u64* launches, ChalState* chal = nullptr, u32* root_dst = nullptr,
                         u32* sample_out = nullptr, u32 n_sample = 0, const u32* pair_layer = nullptr, u64 first = 0, u64 seg = 0,
                         u32 leaf_arity = 2);
int launch_challenger(ChalState* st_dev, const u32* in, u32 n_in, u32* out, u32 n_out, u32 bits, cudaStream_t st, u64* launches);
// observe hash_tree(words) (docs/PROVER_SPEC.md section 2) and sample n_sample elements; tree_scratch: hash_tree_scratch_words(n_words) words
u64 hash_tree_scratch_words(u32 n_words);
int launch_observe_hash_tree(const u32* words, u32 n_words, u32* tree_scratch, ChalState* chal, u32* sample_out, u32 n_sample, cudaStream_t st, u64* launches);
int launch_pow_grind(const ChalState* st_dev, u32 bits, u32* result, cudaStream_t st, u64* launches);

// ---- quotient.cu
struct QuotientArgs {
  const u32* lde;      // [width + aux][B][N] Montgomery, coset-major: row z*N+i is the point of natural index i*B+z; main columns then aux columns
  const u32* publde;   // [pub][B][N] Montgomery, same order: the public columns (range table, program ROM)
  u32* q;              // [4][M], natural order
  u32 log_n, log_blowup;
  const u32* pv;       // device, num_public Montgomery values
  const u32* alpha;    // device, ext4 Montgomery
  const u32* lookup;   // device, lookup challenges z[4], theta[4], then the public I/O transcript's sum sio[4] (launch_io_sum), Montgomery
  const u32* xs;       // [M] coset-major: x = shift*w^(i*B+z) at z*N+i
  const u32* dinv;     // [M] 1/(x - 1), same order
  u32* apow_scratch;   // [K][4] device scratch for alpha powers
  // row segment (one proof sharded over several GPUs): only the points j in [seg_j0, seg_j0 + 2^seg_log_nj) of every coset, i.e.
  // the natural indices [seg_j0 << log_blowup, ...); seg_log_nj = 0xffffffff: all rows.  Reads one halo row (j + 1 mod N).
  u32 seg_log_nj = 0xffffffffu; u64 seg_j0 = 0;
  // plane-sharded quotient commit: plane k is stored into q_plane[k] (the matrix of the rank that will transform that plane,
  // possibly peer memory over NVLink) instead of q; nullptr = q
  u32* q_plane[4] = {nullptr, nullptr, nullptr, nullptr};
};
int launch_quotient(const QuotientArgs& a, cudaStream_t st, u64* launches);        // core profile
int launch_quotient_full(const QuotientArgs& a, cudaStream_t st, u64* launches);   // full profile (quotient.cu compiled with ZKIR_PROFILE_FULL)
int launch_domain_tables(u32* xs, u32* dinv, u32 log_n, u32 log_b, u32 shift_canon, cudaStream_t st, u64* launches);

// ---- trace_expand.cu: raw interpreter rows -> AIR columns on the device
struct ExpandArgs {
  const u64* pcs;        // [T]
  const u32* ins;        // [T]
  const u64* regs;       // [T][16] pre-state
  u64 T, N;              // live rows, padded rows (power of two)
  u64 final_regs[16];    // state after the last instruction (padding rows, last READ value)
  u64 final_pc;
  u32* cols;             // out [88][N], canonical
  u64* err;              // out: min over offending rows of (row << 8 | reason); ~0 = none
  u32 col_lo = 0, col_hi = 0xffffffffu;   // only columns [col_lo, col_hi) are written (the multiplicity columns always are)
  u32 n_code = 0;        // program length: rows whose pc is outside [0x1000, 0x1000 + 4 n_code) are errors
};
int launch_trace_expand(const ExpandArgs& a, cudaStream_t st, u64* launches);
// full profile (trace_expand.cu compiled with ZKIR_PROFILE_FULL): the same rows plus, for every load / store row, the aligned 8-byte word
// before the access and the timestamp of its previous access (from the host's replay of the run's memory: host/pack.cc zkir_mem_replay_full).
// Writes every per-row column of the full table and the histogram columns; the boundary cells of the memory argument (final values of
// image / RAM words: a few rows) are uploaded by the caller afterwards.
struct ExpandFullArgs {
  ExpandArgs rows;
  const u64* old_word;   // [T] little-endian word before the access (memory rows only)
  const u32* prev_ts;    // [T]
};
int launch_trace_expand_full(const ExpandFullArgs& a, cudaStream_t st, u64* launches);
// ... and from the interpreter's register write log + memory log (16 + 12 B per row; zkir_vm_run_writelog_mem_cb): registers by the
// last-writer scan (launch_wl_prefix, then pass 3 inside the kernel), memory cells from the logged word / previous timestamp
struct WlFullArgs {
  const u32* pcs;        // [T]
  const u32* ins;        // [T]
  const u64* wlog;       // [T]
  const u64* old_word;   // [T] (memory rows only)
  const u32* prev_ts;    // [T]
  u64 T, N;
  u64 final_pc;
  int* chunk_prev;       // scratch, trace_expand_wl_scratch_ints(N) ints
  u32* cols;
  u64* err;
  u32 n_code = 0;
};
int launch_wl_prefix(const u64* wlog, u64 T, u64 N, int* chunk_prev, cudaStream_t st, u64* launches);   // core build of trace_expand.cu
int launch_trace_expand_wl_full(const WlFullArgs& a, cudaStream_t st, u64* launches);
int launch_add_u32(u32* dst, const u32* src, u64 n, cudaStream_t st, u64* launches);   // dst[i] += src[i]
struct WlArgs {          // register write log instead of full rows (trace_expand.cu)
  const u32* pcs;        // [T]
  const u32* ins;        // [T]
  const u64* wlog;       // [T]: (k << 56) | value when the row changed register k, else 0
  u64 T, N;
  u64 final_pc;
  int* chunk_prev;       // scratch, trace_expand_wl_scratch_ints(N) ints
  u32* cols;
  u64* err;
  u32 col_lo = 0, col_hi = 0xffffffffu;
  u32 n_code = 0;
};
u64 trace_expand_wl_scratch_ints(u64 N);

// ---- aux_gen.cu: the LogUp aux columns (three helper sums + running sum, ext4 as 4 base columns each) of a trace
struct AuxArgs {
  const u32* trace;      // [88][N] canonical main columns
  const u32* pub;        // [4][N] canonical public columns
  const u32* lookup;     // device: z[4], theta[4], sio[4] (the sum the fractions of all rows must add up to), Montgomery
  u32* aux;              // out [16][N] canonical
  u32 log_n;
  E4* row_tot;           // scratch [N]: sum of the fractions of each row
  E4* blk_tot;           // scratch [aux_gen_blocks(N)]
  u64* err;              // out (may be null): (N-1) << 8 | 8 if the lookups do not balance
};
u64 aux_gen_blocks(u64 N);
// lookup[8..12) <- S_io = sum over the public I/O events (4 words each: clk, kind, lo, hi; canonical, device) of 1 / (z - fingerprint)
int launch_io_sum(const u32* events, u32 n_events, u32* lookup, cudaStream_t st, u64* launches);
int launch_aux_gen(const AuxArgs& a, cudaStream_t st, u64* launches);        // core profile
int launch_aux_gen_full(const AuxArgs& a, cudaStream_t st, u64* launches);   // full profile
int launch_aux_scan(const AuxArgs& a, u32* phi_cols, cudaStream_t st, u64* launches);   // running sum of row_tot into the four phi columns
int launch_trace_expand_wl(const WlArgs& a, cudaStream_t st, u64* launches);

// ---- peer.cu: rows of the column-sharded LDE stored straight into the peers' matrices (NVLink peer memory)
int launch_lde_scatter(const u32* lde, const PeerPtrs& peers, u32 me, u32 G, u32 c_lo, u32 n_cols, u64 N, u32 B, u64 nj, cudaStream_t st,
                       u64* launches);

// ---- stark.cu (openings, DEEP combination, FRI fold, queries, misc)
int launch_map(u32* dst, const u32* src, u64 n, int to_mont, cudaStream_t st, u64* launches);
// out[pos] = (base_ext * mul_const)^(k(pos)), k = coefficient index of memory position pos under `plan` (nd = 1: natural)
int launch_ext_powers(const u32* base_ext, u32 mul_const, const FastPlan& plan, E4* out, cudaStream_t st, u64* launches);
// out1[k] = sum_j coef[k][j]*U1[j], out2[k] = sum_j coef[k][j]*U2[j]
int launch_open(const u32* coef, u64 col_stride, u32 n_cols, u64 n, const E4* U1, const E4* U2, E4* out1, E4* out2,
                E4* partial_scratch, cudaStream_t st, u64* launches);
struct DeepArgs {
  const u32* lde; u64 M; u32 width;       // trace LDE [width][M], coset-major rows
  const u32* qlde; u32 qwidth;            // quotient LDE [8][M], coset-major rows
  u32 log_n, log_b;
  const u32* xs;                          // [M] coset-major
  const u32* zeta; u32 g_mont;            // device ext4; generator of H_N (Montgomery)
  const u32* alpha_fri;                   // device ext4
  const E4* open_t; const E4* open_tg; const E4* open_q;  // device openings
  E4* afp_scratch;                        // [2*width+qwidth+3] scratch
  E4* out;                                // [M], natural order
  u32 seg_log_nj = 0xffffffffu; u64 seg_j0 = 0;   // row segment as in QuotientArgs
};
int launch_deep(const DeepArgs& a, cudaStream_t st, u64* launches);
// square_beta = s: fold with beta^(2^s) (the later half-folds of a fold-by-4/8 round)
int launch_fri_fold(const E4* in, E4* out, u64 h, const u32* beta_dev, const u32* inv_w_table, u32 tw_stride, u32 c_mont,
                    cudaStream_t st, u64* launches, int square_beta = 0);
struct QueryArgs {
  const u32* indices;   // device [num_queries], canonical
  u32 num_queries, log_m, width, log_n;   // lde / qlde rows are coset-major, trees and layers natural
  u32 log_lr = 0;                          // a matrix leaf holds 2^log_lr consecutive natural rows; all of them are opened
  u32 aux_width = 0; const u32* atree = nullptr; u32 atree_sl = 0;   // aux columns = columns [width, width + aux_width) of lde, own tree
  const u32* lde; const u32* ttree;
  const u32* qlde; const u32* qtree;
  const E4* const* layers;        // device array of layer pointers, indexed by fold LEVEL (layer length M >> level)
  const u32* const* ltrees;       // device array of layer trees, same index (only committed levels are used)
  u32 fold8_rounds = 0, last_log_arity = 0, fri_rounds = 0;   // log_n / 3 rounds fold by 8, one more by 2^(log_n mod 3) (docs/PROVER_SPEC.md 4.6)
  u32* out;                       // proof words at the start of the query section
  u32 words_per_query;
  // one proof sharded over several GPUs: this context owns the leaf segments [shard_lo, shard_hi); the lowest *_sl levels of a
  // tree exist only on the owner of the leaf (segment = 2^sl leaves).  Pieces this context does not own are written as 0.
  u32 shard_lo = 0, shard_hi = 1, ttree_sl = 0, qtree_sl = 0;
  u32 lde_sl = 0;       // != 0: the trace LDE rows themselves are row-sharded with segments of 2^lde_sl leaves
  u32 qlde_sl = 0;      // same for the quotient LDE rows
  u32 layer_sl[32] = {0};
};
// LA half-folds of one committed FRI round in one launch: in = layer of n values, out = layer of n >> la values (la = 1..3);
// c_mont[s] = 1 / (2 * shift^(2^s)) in Montgomery form, tw_stride = 1 << (level of `in`)
int launch_fri_fold_multi(const E4* in, E4* out, u64 n, u32 la, const u32* beta_dev, const u32* inv_w_table, u32 tw_stride, const u32 c_mont[3],
                          cudaStream_t st, u64* launches);
int launch_queries(const QueryArgs& a, cudaStream_t st, u64* launches);

}  // namespace zkir
