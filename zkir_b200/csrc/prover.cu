// C ABI + orchestration of the trace->proof path on B200s (include/zkir_b200.h).
//
// Sits where a Rust `zkir_runtime::prove()` would call into a prover; the reference has neither (its runtime API
// ends at run()/VM::run, zkir-runtime/src/lib.rs:29-62, vm.rs:54-78).  Protocol: docs/PROVER_SPEC.md.
// Everything between the H2D copy of the trace and the D2H copy of the proof is a fixed sequence of kernel
// launches on one stream: the Fiat-Shamir challenger, the PoW grinder and the query sampler run on the device, so
// the host never synchronises inside a proof.  Three execution shapes share prove_resident():
//   * one context, ordinary launches (any size);
//   * one context, the whole sequence captured once into a CUDA graph and replayed (proofs of up to 2^12 rows, which are
//     bound by launch and permutation latency: zkir_b200_prove, zkir_b200_prove_batch);
//   * several contexts linked by zkir_b200_comm_init, ONE proof sharded over their GPUs: columns for the
//     per-column transforms, row segments for the per-row sweeps, NVLink peer stores fused into the producing kernels for the
//     bulk exchanges (ntt_fast.cu TileParams::peer, quotient.cu q_plane), NCCL (comm.cu, bound with dlopen) for segment roots,
//     openings, FRI layer 0, query pieces and the barriers.  zkir_b200_emulate_shards runs that control flow on one GPU.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <string>
#include <vector>
#include <map>
#include <thread>
#include "zkir_b200.h"
#include "bb.cuh"
#include "kernels.h"
#include "comm.h"
#include "constants_generated.h"
#include "air_profiles_generated.h"
#include "air_columns.h"   // the device converter and the write-log path serve the core profile
#include "air_pack.h"
#include "proof_layout.h"

using namespace zkir;
namespace zkir { u64 open_scratch_elems(u32 n_cols, u64 n); }

std::string& zkir_host_error();
#define g_last_error zkir_host_error()

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                              \
      return e__ == cudaErrorMemoryAllocation ? ZKIR_ERR_OOM : ZKIR_ERR_CUDA;                      \
    }                                                                                              \
  } while (0)
#define RC(call)                                                                                   \
  do {                                                                                             \
    int r__ = (call);                                                                              \
    if (r__ != 0) {                                                                                \
      if (ctx->err.empty()) ctx->err = std::string(#call) + " failed";                             \
      cudaError_t le = cudaGetLastError();                                                         \
      if (le != cudaSuccess) ctx->err += std::string(": ") + cudaGetErrorString(le);               \
      return r__ == -4 ? ZKIR_ERR_OOM : (r__ == -1 ? ZKIR_ERR_ARG : ZKIR_ERR_CUDA);                \
    }                                                                                              \
  } while (0)

static u32 hpow(u32 a, u64 e) { u64 r = 1, b = a; while (e) { if (e & 1) r = r * b % BB_P; b = b * b % BB_P; e >>= 1; } return (u32)r; }
static u32 hinv(u32 a) { return hpow(a, BB_P - 2); }
static u32 hmul(u32 a, u32 b) { return (u32)((u64)a * b % BB_P); }

extern "C" void zkir_program_digest(const uint32_t* code, size_t n_code, uint32_t digest8[8]);   // host/verify.cc
extern "C" void zkir_io_digest(const uint32_t* io_events, size_t n_io, uint32_t digest8[8]);
static const u32 HDR_WORDS = 7;   // transcript header: log_n, W, aux width, log_blowup, num_queries, pow_bits, num_public

enum { PEER_LDE = 0, PEER_Q = 1, PEER_QLDE = 2, PEER_BUFS = 3 };
#define PEER_REC_WORDS 60
struct Workspace {
  u32 log_n = 0, log_blowup = 0, width = 0, nq = 0;
  bool valid = false;
  std::vector<void*> allocs;
  u32 *trace = nullptr, *coef = nullptr, *lde = nullptr, *ttree = nullptr;   // coef / lde: main columns [0, W) then aux columns [W, W + AW)
  u32 *aux = nullptr, *atree = nullptr;            // aux columns [AW][N] (canonical), their Merkle tree
  u32 *pub = nullptr, *publde = nullptr;           // public columns [PW][N] canonical and their LDE [PW][M] (never committed)
  u64 pub_version = 0;                             // program version publde was built for (0 = none)
  E4 *aux_row_tot = nullptr, *aux_blk_tot = nullptr;
  u32 *q = nullptr, *qcoef = nullptr, *qlde = nullptr, *qtree = nullptr;
  u32 *xs = nullptr, *dinv = nullptr;
  E4 *U1 = nullptr, *U2 = nullptr, *U1q = nullptr, *open_scratch = nullptr, *dummy_open = nullptr;
  bool fast = false;            // register-tile NTT path (log_n >= 8): digit-reversed coefficients
  FastPlan plan_n, plan_m, plan_chunk;
  E4* layers = nullptr;      // all FRI layers back to back: M + M/2 + ... + B
  u32* ltrees = nullptr;     // all layer trees back to back
  std::vector<E4*> h_layers; std::vector<u32*> h_ltrees;
  E4** d_layers = nullptr; u32** d_ltrees = nullptr;
  ChalState* chal = nullptr;
  u32* chal_buf = nullptr;   // alpha[4] zeta[4] alpha_fri[4] betas[R][4] pow_raw[1] pow_sample[1] lookup z[4] theta[4] sio[4] hdr_mont[7+np+16]
  u32* indices = nullptr;
  u32* apow = nullptr; E4* afp = nullptr;
  u32* otree = nullptr;      // scratch of the tree hash of the opened values
  u32* proof = nullptr;      // device proof words
  u32* h_proof = nullptr;    // pinned
  u32* h_stage = nullptr;    // pinned staging for header words
  // sharded proofs: every rank's LDE matrix as mapped into this process (CUDA IPC / peer access), see peer.cu
  PeerPtrs peers[3]; bool peers_valid = false;   // PEER_LDE, PEER_Q, PEER_QLDE
  std::vector<void*> ipc_opened;
  // small proofs are launch-latency bound (about 100 dependent launches): after one ordinary proof of a shape (which fills every
  // table cache) the whole H2D -> kernels -> D2H sequence is captured once into a CUDA graph and replayed for later proofs
  cudaGraphExec_t gexec = nullptr;
  u32 proofs_done = 0; bool graph_failed = false, graph_run = false;
  u32 graph_pow_bits = 0;         // the one proving parameter that is not part of the workspace shape
  size_t graph_n_io = 0; const u32* graph_d_io = nullptr;   // the I/O transcript's length and device buffer are baked into the captured launches
  u64 graph_launches = 0;         // kernels in one proof of this shape (the launch counter advances by this per replay)
  u32* h_trace_stage = nullptr;   // pinned, [W][N]: the fixed source address of the graph's H2D node
  u32* xchg = nullptr;       // [ZKIR_MAX_SHARDS][PEER_REC_WORDS] words: handle exchange; word 0 doubles as the barrier token
};

struct zkir_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  NttTables* tables = nullptr;
  FastNtt* fast = nullptr;
  u64 launches = 0;
  std::string err;
  u32* ntt_tmp = nullptr; u64 ntt_tmp_words = 0;
  u32* scratch[2] = {nullptr, nullptr}; size_t scratch_bytes[2] = {0, 0};
  void* rows_dev = nullptr; size_t rows_bytes = 0;   // staging for raw interpreter rows (prove_rows)
  void* memlog_pinned = nullptr; u64 memlog_capacity = 0;   // pinned write log + memory log of prove_program (full profile), 28 B per cycle
  u32* full_stage = nullptr; size_t full_stage_bytes = 0;   // pinned host staging of the full profile (prove_rows): memory replay arrays, boundary cells
  // zkir_b200_prove_program: pinned write log the interpreter records into, and the stream its chunks are uploaded on
  void* log_pinned = nullptr; u64 log_capacity = 0;
  cudaStream_t copy_stream = nullptr; cudaEvent_t copy_done = nullptr;
  u64* d_err = nullptr; u64* h_err = nullptr;   // [2]: converter error, aux (lookup balance) error; ~0 = none
  std::vector<zkir_ctx*> workers;                    // extra contexts of the same device for prove_batch
  std::vector<u32> code;                             // the program (zkir_b200_set_program): ROM of the lookup argument
  u32 code_digest[8] = {0};                          // its transcript digest
  u64 program_version = 0;                           // bumped by every set_program; workspaces rebuild their ROM columns lazily
  std::vector<u32> io;                               // public I/O transcript of the next proof (zkir_b200_set_io), 4 words per event
  u32 io_digest[8] = {0};                            // its transcript digest (of the empty list until set)
  bool io_digest_valid = false;
  u32* d_io = nullptr; size_t d_io_cap = 0;          // device copy
  // one proof sharded over `shards` GPUs (zkir_b200_comm_init): this context computes the Merkle leaf segments
  // [shard_lo, shard_hi) -- its own rank with a communicator, all of them when the shards are emulated on one GPU (tests)
  Comm* comm = nullptr;
  u32 shards = 1, shard_lo = 0, shard_hi = 1;
  u64 shard_min_seg = 4096;                          // trees with fewer leaves per shard are built whole on every rank
  Workspace ws;
  cudaEvent_t ev[ZKIR_STAGE_COUNT + 1];
  cudaEvent_t tev[2];
  float stage_ms[ZKIR_STAGE_COUNT] = {0};
  bool have_stage = false;
};

static int peers_barrier(zkir_ctx* ctx);
static void peers_close(zkir_ctx* ctx) {
  Workspace& w = ctx->ws;
  for (void* p : w.ipc_opened) cudaIpcCloseMemHandle(p);
  w.ipc_opened.clear();
  w.peers_valid = false;
}
static void ws_free(zkir_ctx* ctx) {
  Workspace& w = ctx->ws;
  peers_close(ctx);
  for (void* p : w.allocs) cudaFree(p);
  if (w.gexec) cudaGraphExecDestroy(w.gexec);
  if (w.h_trace_stage) cudaFreeHost(w.h_trace_stage);
  if (w.h_proof) cudaFreeHost(w.h_proof);
  if (w.h_stage) cudaFreeHost(w.h_stage);
  w = Workspace();
}
template <class T>
static int ws_alloc(zkir_ctx* ctx, T** p, size_t count) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return ZKIR_ERR_OOM; }
  ctx->ws.allocs.push_back(d);
  *p = reinterpret_cast<T*>(d);
  return 0;
}
static int ensure_ntt_tmp(zkir_ctx* ctx, u64 n) {
  // batch scratch for multi-pass NTTs: default 32 MiB so that a batch stays L2-resident between passes
  u64 want = 8ull << 20;  // words
  const char* env = getenv("ZKIR_NTT_BATCH_MB");
  if (env) want = (u64)atoi(env) * (1ull << 18);
  if (want < n) want = n;
  if (ctx->ntt_tmp_words >= want) return 0;
  if (ctx->ntt_tmp) cudaFree(ctx->ntt_tmp);
  ctx->ntt_tmp = nullptr; ctx->ntt_tmp_words = 0;
  CU(cudaMalloc(&ctx->ntt_tmp, want * 4));
  ctx->ntt_tmp_words = want;
  return 0;
}

// persistent scratch for the per-kernel entry points (cudaMallocAsync would return the pool to the OS at every sync)
static int ensure_scratch(zkir_ctx* ctx, int which, size_t bytes, u32** out) {
  if (ctx->scratch_bytes[which] < bytes) {
    if (ctx->scratch[which]) cudaFree(ctx->scratch[which]);
    ctx->scratch[which] = nullptr; ctx->scratch_bytes[which] = 0;
    CU(cudaMalloc(&ctx->scratch[which], bytes));
    ctx->scratch_bytes[which] = bytes;
  }
  *out = ctx->scratch[which];
  return 0;
}

// NTT path: the fast kernels need every digit of the plans (trace, LDE domain, quotient chunks) to span >= 16 lanes
static bool fast_path_ok(u32 log_n, u32 log_blowup, FastPlan* plan_n, FastPlan* plan_m) {
  return !getenv("ZKIR_FORCE_GENERIC_NTT") && fast_plan((int)log_n, plan_n) && fast_plan((int)(log_n + log_blowup), plan_m) &&
         plan_m->d[plan_m->nd - 1] - (int)log_blowup >= 4;
}

static int ws_prepare(zkir_ctx* ctx, const zkir_params* p, u32 log_n) {
  Workspace& w = ctx->ws;
  if (w.valid && w.log_n == log_n && w.log_blowup == p->log_blowup && w.width == p->width && w.nq == p->num_queries) return 0;
  if (ctx->comm && w.valid) {
    // Sharded mode: other ranks may still have this workspace mapped (cudaIpcOpenMemHandle); freeing exported memory before the
    // importers closed it is undefined.  A shape change is collective, so: everyone closes its mappings, one barrier, then free.
    peers_close(ctx);
    int brc = peers_barrier(ctx); if (brc) return brc;
    CU(cudaStreamSynchronize(ctx->stream));
  }
  ws_free(ctx);
  const u64 N = 1ull << log_n, M = N << p->log_blowup, W = p->width;
  const u32 R = log_n;
  const Layout L = make_layout(p, log_n);
  int rc;
#define A(ptr, cnt) if ((rc = ws_alloc(ctx, &w.ptr, (cnt))) != 0) return rc;
  // NTT path: the fast kernels need every digit of the three plans to span >= 16 lanes
  w.fast = fast_path_ok(log_n, p->log_blowup, &w.plan_n, &w.plan_m);
  if (!w.fast) { ctx->err = "the register-tile NTT plan does not cover this shape (ZKIR_FORCE_GENERIC_NTT only applies to the per-kernel entry points)"; return ZKIR_ERR_ARG; }
  w.plan_chunk = w.plan_m;
  w.plan_chunk.log_n = (int)log_n;
  w.plan_chunk.d[w.plan_m.nd - 1] -= (int)p->log_blowup;
  const u64 AW = profile_aux_width(p->width), PW = profile_pub_width(p->width), WA = W + AW;
  A(trace, W * N) A(coef, WA * N) A(lde, WA * M) A(ttree, (2 * M - 1) * 8)
  A(aux, (u64)AW * N) A(atree, (2 * M - 1) * 8) A(pub, (u64)PW * N) A(publde, (u64)PW * M)
  A(aux_row_tot, N) A(aux_blk_tot, aux_gen_blocks(N))
  A(q, 4 * M) A(qlde, QW * M) A(qtree, (2 * M - 1) * 8)
  A(qcoef, QW * N)
  A(xs, M) A(dinv, M)
  A(U1, N) A(U2, N) A(U1q, N) A(open_scratch, open_scratch_elems((u32)WA, N)) A(dummy_open, WA)
  A(layers, 2 * M) A(ltrees, 2 * M * 8)
  A(d_layers, R + 1) A(d_ltrees, R + 1)
  A(chal, 1) A(chal_buf, 12 + 4 * R + 2 + 12 + HDR_WORDS + p->num_public + 16) A(indices, p->num_queries + 1)
  A(apow, 4 * ZKIR_PROFILE_MAX_CONSTRAINTS) A(afp, WA + 5)
  A(proof, L.total + 4) A(xchg, ZKIR_MAX_SHARDS * PEER_REC_WORDS) A(otree, hash_tree_scratch_words((u32)(2 * WA + QW) * 4))
#undef A
  w.proof += (4 - (L.open_t & 3)) & 3;   // the opened values are read and written as 16-byte ext4 elements: align that section
  CU(cudaMallocHost(&w.h_proof, L.total * 4));
  CU(cudaMallocHost(&w.h_stage, (8 + HDR_WORDS + 2 * p->num_public + 16) * 4));
  // layer pointers
  w.h_layers.resize(R + 1); w.h_ltrees.resize(R + 1);
  E4* lp = w.layers; u32* tp = w.ltrees;
  for (u32 r = 0; r <= R; r++) {
    w.h_layers[r] = lp; w.h_ltrees[r] = tp;
    const u64 n = M >> r;
    lp += n; tp += n * 8;  // tree of n/2 leaves needs (n-1)*8 words
  }
  CU(cudaMemcpyAsync(w.d_layers, w.h_layers.data(), (R + 1) * sizeof(E4*), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(w.d_ltrees, w.h_ltrees.data(), (R + 1) * sizeof(u32*), cudaMemcpyHostToDevice, ctx->stream));
  RC(launch_domain_tables(w.xs, w.dinv, log_n, p->log_blowup, ZKIR_BB_GEN, ctx->stream, &ctx->launches));
  CU(cudaStreamSynchronize(ctx->stream));
  w.log_n = log_n; w.log_blowup = p->log_blowup; w.width = p->width; w.nq = p->num_queries; w.valid = true;
  return 0;
}

static int check_params(zkir_ctx* ctx, const zkir_params* p, u32 log_n) {
  if (!p || !profile_known(p->width) || p->num_public != ZKIR_NUM_PUBLIC_VALUES || p->log_blowup < 1 || p->log_blowup > 4 ||
      log_n < ZKIR_RANGE_BITS || log_n + p->log_blowup > 27 || p->pow_bits > 30 || p->num_queries > 4096) {
    ctx->err = "bad params: need width=88, num_public=5, 1<=log_blowup<=4, 10<=log_n (the range table occupies 1024 trace rows), log_n+log_blowup<=27, pow_bits<=30";
    return ZKIR_ERR_ARG;
  }
  if (ctx->program_version == 0) { ctx->err = "no program: call zkir_b200_set_program first (the proof binds the executed instructions to it)"; return ZKIR_ERR_ARG; }
  if (ctx->code.size() > (1ull << log_n)) { ctx->err = "the program has more instructions than the trace has rows: raise log_n"; return ZKIR_ERR_ARG; }
  return 0;
}

// One proof over `shards` GPUs: the columns are cut into `shards` contiguous ranges for the per-column work
// (LDE, openings), the rows into `shards` contiguous ranges of the trace domain (points j in [g*N/G, (g+1)*N/G) of every
// coset = the natural-order leaves [g*M/G, (g+1)*M/G)) for the per-row work (leaf hashing, quotient, DEEP, queries).
struct ShardPlan {
  bool on = false;
  u32 G = 1, lo = 0, hi = 1, log_nj = 0;
  u64 nj = 0;          // points per coset and shard
  u32 cols_per = 0, W = 0;
  u32 AW = 0;          // aux columns of the profile
  u32 acols_per = 0;   // aux columns per rank (column indices W .. W + AW of the coefficient / LDE matrices)
  u32 a_lo(u32 g) const { const u32 c = g * acols_per; return W + (c < AW ? c : AW); }
  u32 a_hi(u32 g) const { const u32 c = (g + 1) * acols_per; return W + (c < AW ? c : AW); }
  u32 planes_per = 1;  // quotient planes (of 4) per rank; ranks beyond 4 / planes_per own none
  u32 p_lo(u32 g) const { const u32 c = g * planes_per; return c < 4 ? c : 4; }
  u32 p_hi(u32 g) const { const u32 c = (g + 1) * planes_per; return c < 4 ? c : 4; }
  u32 c_lo(u32 g) const { const u32 c = g * cols_per; return c < W ? c : W; }
  u32 c_hi(u32 g) const { const u32 c = (g + 1) * cols_per; return c < W ? c : W; }
};

// the partition itself: pure arithmetic, also exported as zkir_b200_shard_plan for host-side tests
static ShardPlan shard_plan_for(u32 G, u32 W, u32 log_n, u32 log_blowup, bool fast, u64 min_seg) {
  const u64 N = 1ull << log_n, M = N << log_blowup;
  ShardPlan sp;
  sp.G = G; sp.W = W;
  if (G > 1 && fast && N >= G && M / G >= min_seg && M / G >= 2) {
    sp.on = true;
    sp.nj = N / G;
    while ((1ull << sp.log_nj) < sp.nj) sp.log_nj++;
    sp.cols_per = (W + G - 1) / G;
    sp.AW = profile_aux_width(W);
    sp.acols_per = (sp.AW + G - 1) / G;
    sp.planes_per = (4 + G - 1) / G;
  }
  return sp;
}
// the plan of one proof; needs the workspace of that shape (ws_prepare) for `fast`
static ShardPlan make_shard_plan(const zkir_ctx* ctx, const zkir_params* p, u32 log_n) {
  ShardPlan sp = shard_plan_for(ctx->shards, p->width, log_n, p->log_blowup, ctx->ws.fast, ctx->shard_min_seg);
  sp.lo = ctx->shard_lo; sp.hi = ctx->shard_hi;
  return sp;
}
// columns of the trace this context has to materialise: all of them, or -- sharded proof with a communicator -- its own share
static void trace_col_range(const zkir_ctx* ctx, const zkir_params* p, u32 log_n, u32* lo, u32* hi) {
  // AIR v2: the aux (LogUp) columns of a row depend on the whole row, so every rank materialises all main columns (the converter
  // is HBM-write bound, 0.5 ms at 2^20 rows); only the transforms are column-sharded
  *lo = 0; *hi = p->width;
  (void)ctx; (void)log_n;
}

// After the column-sharded LDE every rank holds whole columns [c_lo, c_hi) of the coset-major matrix [W][B][N]; this moves, for
// every column, the rows each OTHER rank needs (its points j0..j0+nj of every coset plus the one halo row the quotient reads
// as "next row") into the same place of that rank's matrix: (G-1)/G of a rank's W/G columns leave it, i.e. about 1/G of the matrix
// per rank crosses NVLink, instead of the whole matrix for an all-gather.
static int exchange_lde_rows(zkir_ctx* ctx, const ShardPlan& sp, u32* lde, u64 N, u32 B) {
  const u64 M = N * B;
  const int me = comm_rank(ctx->comm);
  std::vector<P2POp> ops;
  auto pieces = [&](u32 c, int owner_of_rows, int peer, int is_send) {
    const u64 j0 = (u64)owner_of_rows * sp.nj;
    const bool wrap = j0 + sp.nj == N;      // the last shard's halo row is point 0
    for (u32 z = 0; z < B; z++) {
      u32* base = lde + (u64)c * M + (u64)z * N;
      ops.push_back({peer, is_send, base + j0, (size_t)(wrap ? sp.nj : sp.nj + 1)});
      if (wrap) ops.push_back({peer, is_send, base, 1});
    }
  };
  for (int peer = 0; peer < (int)sp.G; peer++) {
    if (peer == me) continue;
    for (u32 c = sp.c_lo(me); c < sp.c_hi(me); c++) pieces(c, peer, peer, 1);   // my columns, the peer's rows
    for (u32 c = sp.c_lo(peer); c < sp.c_hi(peer); c++) pieces(c, me, peer, 0); // the peer's columns, my rows
  }
  if (comm_exchange_u32(ctx->comm, ops.data(), ops.size(), ctx->stream, &ctx->err) != 0) return ZKIR_ERR_NCCL;
  return 0;
}

// Map every rank's LDE matrix into this process.  Collective (one all-gather of 96-byte records); runs once per workspace shape.
struct PeerRec { u64 pid; u32 dev, pad; u64 ptr[PEER_BUFS]; cudaIpcMemHandle_t handle[PEER_BUFS]; u64 pad2; };
static_assert(ZKIR_NUM_PUBLIC_VALUES <= 8, "zkir_b200_quotient scratch layout");
static_assert(sizeof(PeerRec) == 4 * PEER_REC_WORDS, "PeerRec layout");
static int peers_open(zkir_ctx* ctx) {
  Workspace& w = ctx->ws;
  peers_close(ctx);
  const int G = comm_world(ctx->comm), me = comm_rank(ctx->comm);
  u32* bufs[PEER_BUFS] = {w.lde, w.q, w.qlde};
  PeerRec mine;
  memset(&mine, 0, sizeof(mine));
  mine.pid = (u64)getpid(); mine.dev = (u32)ctx->device;
  for (int b = 0; b < PEER_BUFS; b++) {
    mine.ptr[b] = (u64)(uintptr_t)bufs[b];
    CU(cudaIpcGetMemHandle(&mine.handle[b], bufs[b]));
  }
  std::vector<PeerRec> all(G);
  CU(cudaMemcpyAsync(w.xchg + (size_t)me * PEER_REC_WORDS, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
  if (comm_all_gather_u32(ctx->comm, w.xchg, PEER_REC_WORDS, ctx->stream, &ctx->err) != 0) return ZKIR_ERR_NCCL;
  CU(cudaMemcpyAsync(all.data(), w.xchg, sizeof(PeerRec) * G, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int t = 0; t < G; t++) {
    if (t == me) { for (int b = 0; b < PEER_BUFS; b++) w.peers[b].p[t] = bufs[b]; continue; }
    if (all[t].pid == mine.pid) {   // another context of this process (threads as ranks): plain peer access
      if ((int)all[t].dev != ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess((int)all[t].dev, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) { ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return ZKIR_ERR_CUDA; }
      }
      for (int b = 0; b < PEER_BUFS; b++) w.peers[b].p[t] = reinterpret_cast<u32*>((uintptr_t)all[t].ptr[b]);
    } else {
      for (int b = 0; b < PEER_BUFS; b++) {
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, all[t].handle[b], cudaIpcMemLazyEnablePeerAccess));
        w.ipc_opened.push_back(ptr);
        w.peers[b].p[t] = reinterpret_cast<u32*>(ptr);
      }
    }
  }
  w.peers_valid = true;
  return 0;
}
// one-word all-reduce: orders every rank's peer-memory stores (kernels before it) before every rank's reads (kernels after it)
static int peers_barrier(zkir_ctx* ctx) {
  return comm_all_reduce_sum_u32(ctx->comm, ctx->ws.xchg, 1, ctx->stream, &ctx->err) != 0 ? ZKIR_ERR_NCCL : 0;
}

// Merkle commitment of one matrix (leaf i = sponge over LDE row i) or one FRI layer (leaf i = hash(f[i] || f[i+h])), followed
// by the Fiat-Shamir step on its root.  With shards > 1 the leaf range is cut into `shards` contiguous segments: this context
// hashes its segments and builds their subtrees in place in the global tree layout, the segment roots are exchanged with ONE
// all-gather of shards * 8 words ("Merkle-root reduction"), and the top log2(shards) levels + the transcript step run
// redundantly on every rank, so every rank continues with the same challenges.  *shard_levels = number of bottom path levels
// that only the owner of a leaf holds (0 = the whole tree is local).
static int commit_tree(zkir_ctx* ctx, const u32* mat, u32 n_cols, u32 log_b, const u32* pair_layer, u32* tree, u64 n_leaves, u32* root_dst,
                       u32* sample_out, u32 n_sample, u32* shard_levels, int mode = -1 /* -1: by size, 0: whole, 1: sharded */, u32 leaf_arity = 2) {
  cudaStream_t st = ctx->stream;
  u64* LC = &ctx->launches;
  Workspace& w = ctx->ws;
  // `mat`: n_leaves = M / 2B leaves of 2*B natural rows each (launch_leaf_hash_rows); log_n_mat = log2 of the points per coset
  const u64 G = ctx->shards, seg = n_leaves / (G ? G : 1);
  u64 min_seg = ctx->shard_min_seg;
  if (min_seg < 2) min_seg = 2;
  u32 log_n_mat = 1;   // N = 2 * n_leaves points per coset
  if (mat) while ((1ull << log_n_mat) < 2 * n_leaves) log_n_mat++;
  const u64 col_stride = mat ? ((1ull << log_n_mat) << log_b) : 0;
  *shard_levels = 0;
  if (G <= 1 || mode == 0 || (mode < 0 && seg < min_seg) || seg < 1) {
    if (mat) RC(launch_leaf_hash_rows(mat, col_stride, n_cols, log_n_mat, log_b, tree, st, LC, 0, n_leaves));
    RC(launch_merkle_levels(tree, n_leaves, st, LC, w.chal, root_dst, sample_out, n_sample, pair_layer, 0, 0, leaf_arity));
    return 0;
  }
  for (u32 g = ctx->shard_lo; g < ctx->shard_hi; g++) {
    if (mat) RC(launch_leaf_hash_rows(mat, col_stride, n_cols, log_n_mat, log_b, tree, st, LC, g * seg, seg));
    RC(launch_merkle_levels(tree, n_leaves, st, LC, nullptr, nullptr, nullptr, 0, pair_layer, g * seg, seg, leaf_arity));
  }
  u32 ls = 0;
  while ((1ull << ls) < seg) ls++;
  u32* roots = tree + 8 * (2 * n_leaves - 2 * (n_leaves >> ls));  // the level that holds the G segment roots
  if (ctx->comm) {
    if (comm_all_gather_u32(ctx->comm, roots, 8, st, &ctx->err) != 0) return ZKIR_ERR_NCCL;
  }
  RC(launch_merkle_levels(roots, G, st, LC, w.chal, root_dst, sample_out, n_sample));
  *shard_levels = ls;
  return 0;
}

// Host staging of one proof: header words + public values (canonical, copied into the proof) and their Montgomery copies (absorbed
// by the transcript).  The only per-proof host data of the device part, so a captured graph of prove_resident is replayed
// after refilling this buffer (and the pinned trace staging).
static int fill_header_stage(zkir_ctx* ctx, const zkir_params* p, u32 log_n, const u32* pv) {
  const u32 np = p->num_public;
  u32* hs = ctx->ws.h_stage;
  hs[0] = PROOF_MAGIC; hs[1] = PROOF_VERSION; hs[2] = log_n; hs[3] = p->width; hs[4] = p->log_blowup; hs[5] = p->num_queries;
  hs[6] = p->pow_bits; hs[7] = np;
  for (u32 i = 0; i < np; i++) { if (pv[i] >= BB_P) { ctx->err = "public value not canonical"; return ZKIR_ERR_ARG; } hs[8 + i] = pv[i]; }
  if (pv[4] > 1 || (pv[4] == 0 && (pv[2] || pv[3]))) { ctx->err = "public values: halted must be 0/1, and a run that did not halt has no exit code"; return ZKIR_ERR_ARG; }
  u32* hm = hs + 8 + np;  // Montgomery copy for the transcript: header, public values, program digest
  const u32 hdr[HDR_WORDS] = {log_n, p->width, profile_aux_width(p->width), p->log_blowup, p->num_queries, p->pow_bits, np};
  for (u32 i = 0; i < HDR_WORDS; i++) hm[i] = bb_to_mont_c(hdr[i]);
  for (u32 i = 0; i < np; i++) hm[HDR_WORDS + i] = bb_to_mont_c(pv[i]);
  for (u32 i = 0; i < 8; i++) hm[HDR_WORDS + np + i] = bb_to_mont_c(ctx->code_digest[i]);
  if (!ctx->io_digest_valid) { zkir_io_digest(ctx->io.data(), ctx->io.size() / 4, ctx->io_digest); ctx->io_digest_valid = true; }
  for (u32 i = 0; i < 8; i++) hm[HDR_WORDS + np + 8 + i] = bb_to_mont_c(ctx->io_digest[i]);
  return 0;
}

// Public columns of the current program for this workspace shape (docs/PROVER_SPEC.md section 3.3): range table 0..1023, ROM pc,
// decoded word, immediate; built on the host (program-sized), uploaded, extended to the LDE coset once per (program, shape).
static int ensure_public_columns(zkir_ctx* ctx, const zkir_params* p, u32 log_n) {
  Workspace& w = ctx->ws;
  if (w.pub_version == ctx->program_version) return 0;
  const u64 N = 1ull << log_n, M = N << p->log_blowup;
  const u32 PW = profile_pub_width(p->width);
  std::vector<u32> h((size_t)PW * N, 0u);
  zkir_public_columns(p->width, log_n, ctx->code.data(), ctx->code.size(), h.data());   // host/public_cols.cc
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(w.pub, h.data(), h.size() * 4, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));   // h is pageable and goes out of scope
  const u32 c0 = hmul(hinv((u32)(N % BB_P)), (u32)((1ull << 32) % BB_P));
  if (w.fast) {
    u32* coef = w.lde;   // scratch: the main LDE region is rewritten by every proof
    RC(fast_intt(ctx->fast, w.plan_n, w.pub, N, coef, N, PW, c0, nullptr, 32, 0, 0, nullptr, 0, st));
    RC(fast_coset_ntt(ctx->fast, w.plan_n, coef, N, w.publde, M, PW, 1u << p->log_blowup, ZKIR_BB_GEN, ZKIR_BB_ROOTS[log_n + p->log_blowup], 1u, st));
  } else {
    ctx->err = "internal: generic NTT path is not available for traces of 2^10 rows and more";
    return ZKIR_ERR_ARG;
  }
  CU(cudaStreamSynchronize(st));
  w.pub_version = ctx->program_version;
  if (w.gexec) { cudaGraphExecDestroy(w.gexec); w.gexec = nullptr; }   // a captured proof sequence embeds nothing of the program, but be safe
  return 0;
}

// the device part of a proof; `trace` = canonical column-major values already on the device
static int prove_resident(zkir_ctx* ctx, const zkir_params* p, u32 log_n, const u32* pv, const u32* trace) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  u64* LC = &ctx->launches;
  const u64 N = 1ull << log_n, M = N << p->log_blowup, W = p->width, AW = profile_aux_width(p->width), WA = W + AW;
  const bool full = profile_is_full(p->width);
  const u32 log_m = log_n + p->log_blowup, R = log_n, np = p->num_public;
  const Layout L = make_layout(p, log_n);
  const u32 shift = ZKIR_BB_GEN;
  const u32 B = 1u << p->log_blowup;
  if (!w.fast) { ctx->err = "internal: the register-tile NTT plan does not cover this shape"; return ZKIR_ERR_ARG; }
  if (w.pub_version != ctx->program_version) { ctx->err = "internal: public columns are stale"; return ZKIR_ERR_ARG; }
  const ShardPlan sp = make_shard_plan(ctx, p, log_n);
  const bool p2p = sp.on && ctx->comm != nullptr;
  static const int lde_mode = !getenv("ZKIR_LDE_EXCHANGE") ? 0 : (!strcmp(getenv("ZKIR_LDE_EXCHANGE"), "nccl") ? 2 : (!strcmp(getenv("ZKIR_LDE_EXCHANGE"), "scatter") ? 1 : 0));
  if (p2p && !w.peers_valid) { int xrc = peers_open(ctx); if (xrc) return xrc; }
  const u32 me = p2p ? (u32)comm_rank(ctx->comm) : 0;
  u32* c_alpha = w.chal_buf, *c_zeta = w.chal_buf + 4, *c_afri = w.chal_buf + 8, *c_betas = w.chal_buf + 12;
  u32* c_pow_raw = w.chal_buf + 12 + 4 * R, *c_pow_sample = c_pow_raw + 1, *c_lookup = c_pow_raw + 2, *c_hdr = c_lookup + 12;

  // ---- header + public values (host staging filled by fill_header_stage: also the per-proof step of a graph replay)
  { int frc = fill_header_stage(ctx, p, log_n, pv); if (frc) return frc; }
  u32* hs = w.h_stage;
  u32* hm = hs + 8 + np;  // Montgomery copy for the transcript
  CU(cudaMemcpyAsync(w.proof, hs, (8 + np) * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(c_hdr, hm, (HDR_WORDS + np + 16) * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(w.chal, 0, sizeof(ChalState), st));
  CU(cudaMemsetAsync(ctx->d_err + 1, 0xff, 8, st));

  // ---- 1. LDE: iNTT (scale by shift^j/N and lift to Montgomery), zero-pad, forward NTT on the coset
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_LDE], st));
  const u32 c0 = hmul(hinv((u32)(N % BB_P)), (u32)((1ull << 32) % BB_P));  // R/N: scales by 1/N and lifts to Montgomery form
  if (sp.on) {
    // column-sharded: this rank transforms only its W/G columns, and the rows go to their owners.  Default: FUSED -- the last pass
    // of the coset NTT stores every output row straight into the owner's matrix over NVLink (ntt_fast.cu, TileParams::peer), no
    // separate copy; ZKIR_LDE_EXCHANGE=scatter: plain LDE + one scatter kernel (peer.cu); =nccl: plain LDE + grouped ncclSend/Recv.
    // A one-word all-reduce is the barrier that orders every rank's peer stores before every rank's reads.
    const bool fuse_t = p2p && lde_mode == 0 && fast_coset_ntt_can_fuse(w.plan_n, sp.G);
    for (u32 g = sp.lo; g < sp.hi; g++) {
      const u32 k0 = sp.c_lo(g), nc = sp.c_hi(g) - k0;
      if (!nc) continue;
      RC(fast_intt(ctx->fast, w.plan_n, trace + (u64)k0 * N, N, w.coef + (u64)k0 * N, N, nc, c0, nullptr, 32, 0, 0, nullptr, 0, st));
      RC(fast_coset_ntt(ctx->fast, w.plan_n, w.coef + (u64)k0 * N, N, w.lde + (u64)k0 * M, M, nc, B, shift, ZKIR_BB_ROOTS[log_m], 1u, st,
                        fuse_t ? &w.peers[PEER_LDE] : nullptr, me, sp.G));
    }
    if (ctx->comm) {
      if (lde_mode == 2) {
        int xrc = exchange_lde_rows(ctx, sp, w.lde, N, B);
        if (xrc) return xrc;
      } else {
        if (!fuse_t) RC(launch_lde_scatter(w.lde, w.peers[PEER_LDE], me, sp.G, sp.c_lo(me), sp.c_hi(me) - sp.c_lo(me), N, B, sp.nj, st, LC));
        int brc = peers_barrier(ctx); if (brc) return brc;
      }
    }
  } else {
    // column batches (ZKIR_LDE_BATCH) are possible, one batch = all columns by default
    u32 cb = (u32)W;  // measured: the passes are integer-pipe bound, larger launches win over L2 residency
    const char* env = getenv("ZKIR_LDE_BATCH");
    if (env && atoi(env) > 0) cb = (u32)atoi(env);
    for (u32 k0 = 0; k0 < W; k0 += cb) {
      const u32 nc = W - k0 < cb ? (u32)(W - k0) : cb;
      RC(fast_intt(ctx->fast, w.plan_n, trace + (u64)k0 * N, N, w.coef + (u64)k0 * N, N, nc, c0, nullptr, 32, 0, 0, nullptr, 0, st));
      RC(fast_coset_ntt(ctx->fast, w.plan_n, w.coef + (u64)k0 * N, N, w.lde + (u64)k0 * M, M, nc, B, shift, ZKIR_BB_ROOTS[log_m], 1u, st));
    }
  }
  // ---- 2. trace commitment
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_TRACE_COMMIT], st));
  u32 t_sl = 0, a_sl = 0, q_sl = 0;
  int crc;
  RC(launch_challenger(w.chal, c_hdr, HDR_WORDS + np + 16, nullptr, 0, 0, st, LC));   // header, public values, program digest, I/O transcript digest
  if ((crc = commit_tree(ctx, w.lde, (u32)W, p->log_blowup, nullptr, w.ttree, M >> L.log_lr, w.proof + L.troot, c_lookup, 8, &t_sl, sp.on ? 1 : 0)) != 0) return crc;  // root -> proof, observe, sample z and theta
  // ---- 2b. LogUp aux columns for the challenges just drawn (aux_gen.cu), their LDE and their own commitment
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_AUX], st));
  {
    RC(launch_io_sum(ctx->d_io, (u32)(ctx->io.size() / 4), c_lookup, st, LC));   // S_io of the public I/O transcript for the challenges just drawn
    AuxArgs aa;
    aa.trace = trace; aa.pub = w.pub; aa.lookup = c_lookup; aa.aux = w.aux; aa.log_n = log_n; aa.row_tot = w.aux_row_tot; aa.blk_tot = w.aux_blk_tot;
    aa.err = ctx->d_err + 1;
    RC(full ? launch_aux_gen_full(aa, st, LC) : launch_aux_gen(aa, st, LC));   // replicated on every rank of a sharded proof (needs whole rows; 16 columns out)
    const u32 c0a = hmul(hinv((u32)(N % BB_P)), (u32)((1ull << 32) % BB_P));
    if (sp.on) {
      const bool fuse_a = p2p && lde_mode == 0 && fast_coset_ntt_can_fuse(w.plan_n, sp.G);
      for (u32 g = sp.lo; g < sp.hi; g++) {
        const u32 k0 = sp.a_lo(g), nc = sp.a_hi(g) - k0;
        if (!nc) continue;
        RC(fast_intt(ctx->fast, w.plan_n, w.aux + (u64)(k0 - W) * N, N, w.coef + (u64)k0 * N, N, nc, c0a, nullptr, 32, 0, 0, nullptr, 0, st));
        RC(fast_coset_ntt(ctx->fast, w.plan_n, w.coef + (u64)k0 * N, N, w.lde + (u64)k0 * M, M, nc, B, shift, ZKIR_BB_ROOTS[log_m], 1u, st,
                          fuse_a ? &w.peers[PEER_LDE] : nullptr, me, sp.G));
      }
      if (p2p) {
        if (lde_mode == 2) { ctx->err = "ZKIR_LDE_EXCHANGE=nccl is not supported with the aux phase"; return ZKIR_ERR_ARG; }
        if (!fuse_a) RC(launch_lde_scatter(w.lde, w.peers[PEER_LDE], me, sp.G, sp.a_lo(me), sp.a_hi(me) - sp.a_lo(me), N, B, sp.nj, st, LC));
        int brc = peers_barrier(ctx); if (brc) return brc;
      }
    } else {
      RC(fast_intt(ctx->fast, w.plan_n, w.aux, N, w.coef + W * N, N, (u32)AW, c0a, nullptr, 32, 0, 0, nullptr, 0, st));
      RC(fast_coset_ntt(ctx->fast, w.plan_n, w.coef + W * N, N, w.lde + W * M, M, (u32)AW, B, shift, ZKIR_BB_ROOTS[log_m], 1u, st));
    }
    if ((crc = commit_tree(ctx, w.lde + W * M, (u32)AW, p->log_blowup, nullptr, w.atree, M >> L.log_lr, w.proof + L.aroot, c_alpha, 4, &a_sl, sp.on ? 1 : 0)) != 0) return crc;  // root -> proof, observe, sample alpha
  }
  // ---- 3. quotient
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_QUOTIENT], st));
  {
    QuotientArgs qa;
    qa.lde = w.lde; qa.publde = w.publde; qa.q = w.q; qa.log_n = log_n; qa.log_blowup = p->log_blowup; qa.pv = c_hdr + HDR_WORDS; qa.alpha = c_alpha; qa.lookup = c_lookup;
    qa.xs = w.xs; qa.dinv = w.dinv; qa.apow_scratch = w.apow;
    if (sp.on) {
      // row-sharded: every rank evaluates its natural-order range of Q and stores plane k straight into the matrix of the rank
      // that transforms that plane (NVLink peer stores from the quotient kernel itself); a barrier completes the planes
      if (p2p) for (u32 k = 0; k < 4; k++) qa.q_plane[k] = w.peers[PEER_Q].p[k / sp.planes_per];
      for (u32 g = sp.lo; g < sp.hi; g++) {
        qa.seg_log_nj = sp.log_nj; qa.seg_j0 = (u64)g * sp.nj;
        RC(full ? launch_quotient_full(qa, st, LC) : launch_quotient(qa, st, LC));
      }
      if (p2p) { int brc = peers_barrier(ctx); if (brc) return brc; }
    } else {
      RC(full ? launch_quotient_full(qa, st, LC) : launch_quotient(qa, st, LC));
    }
  }
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_QUOTIENT_COMMIT], st));
  const u32* qcoef = w.qcoef;
  {
    // inverse transform over the whole coset (digits of plan_m, in place on q), unshift by shift^-k, and let the last pass
    // split the lowest digit into the two degree-<N chunks: column 2*plane + chunk of qcoef, digit-reversed under plan_chunk
    const uint2* unshift = fast_scale_table(ctx->fast, w.plan_m, hinv(shift), 1u);
    if (!unshift) { ctx->err = "table alloc"; return ZKIR_ERR_OOM; }
    const u32 split_log = (u32)w.plan_chunk.d[w.plan_chunk.nd - 1];
    if (sp.on) {
      // plane-sharded: a rank transforms its planes (2 quotient columns each), then the rows go to their owners like the trace LDE
      const bool fuse_q = p2p && lde_mode == 0 && fast_coset_ntt_can_fuse(w.plan_chunk, sp.G);
      for (u32 g = sp.lo; g < sp.hi; g++) {
        const u32 p0 = sp.p_lo(g), npl = sp.p_hi(g) - p0;
        if (!npl) continue;
        RC(fast_intt(ctx->fast, w.plan_m, w.q + (u64)p0 * M, M, w.q + (u64)p0 * M, M, npl, hinv((u32)(M % BB_P)), unshift, split_log, 2, (u32)N,
                     w.qcoef + (u64)p0 * 2 * N, 2 * N, st));
        RC(fast_coset_ntt(ctx->fast, w.plan_chunk, w.qcoef + (u64)p0 * 2 * N, N, w.qlde + (u64)p0 * 2 * M, M, 2 * npl, B, shift, ZKIR_BB_ROOTS[log_m], 1u, st,
                          fuse_q ? &w.peers[PEER_QLDE] : nullptr, me, sp.G));
      }
      if (p2p) {
        if (!fuse_q) RC(launch_lde_scatter(w.qlde, w.peers[PEER_QLDE], me, sp.G, 2 * sp.p_lo(me), 2 * (sp.p_hi(me) - sp.p_lo(me)), N, B, sp.nj, st, LC));
        int brc = peers_barrier(ctx); if (brc) return brc;
      }
    } else {
      RC(fast_intt(ctx->fast, w.plan_m, w.q, M, w.q, M, 4, hinv((u32)(M % BB_P)), unshift, split_log, 2, (u32)N, w.qcoef, 2 * N, st));
      RC(fast_coset_ntt(ctx->fast, w.plan_chunk, w.qcoef, N, w.qlde, M, QW, B, shift, ZKIR_BB_ROOTS[log_m], 1u, st));
    }
  }
  if ((crc = commit_tree(ctx, w.qlde, QW, p->log_blowup, nullptr, w.qtree, M >> L.log_lr, w.proof + L.qroot, c_zeta, 4, &q_sl, sp.on ? 1 : -1)) != 0) return crc;  // root -> proof, observe, sample zeta
  // ---- 4. openings at zeta and g*zeta, evaluated on the shifted coefficients at zeta/shift
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_OPENINGS], st));
  {
    // fast path: plain coefficients, digit-reversed; generic path: coefficients pre-multiplied by shift^k, natural order
    const u32 g = ZKIR_BB_ROOTS[log_n], um = 1u;   // plain coefficients, digit-reversed (register-tile NTT path)
    RC(launch_ext_powers(c_zeta, bb_to_mont_c(um), w.plan_n, w.U1, st, LC));
    RC(launch_ext_powers(c_zeta, bb_to_mont_c(hmul(um, g)), w.plan_n, w.U2, st, LC));
    RC(launch_ext_powers(c_zeta, bb_to_mont_c(um), w.plan_chunk, w.U1q, st, LC));
    E4* ot = reinterpret_cast<E4*>(w.proof + L.open_t);
    E4* otg = reinterpret_cast<E4*>(w.proof + L.open_tg);
    E4* oq = reinterpret_cast<E4*>(w.proof + L.open_q);
    if (sp.on) {
      // column-sharded like the LDE (a rank holds the coefficients of its own columns only); disjoint pieces merged by all-reduce
      if (ctx->comm) CU(cudaMemsetAsync(w.proof + L.open_t, 0, (2 * WA + QW) * 16, st));
      for (u32 g = sp.lo; g < sp.hi; g++) {
        const u32 k0 = sp.c_lo(g), nc = sp.c_hi(g) - k0;
        if (nc) RC(launch_open(w.coef + (u64)k0 * N, N, nc, N, w.U1, w.U2, ot + k0, otg + k0, w.open_scratch, st, LC));
        const u32 a0 = sp.a_lo(g), na = sp.a_hi(g) - a0;
        if (na) RC(launch_open(w.coef + (u64)a0 * N, N, na, N, w.U1, w.U2, ot + a0, otg + a0, w.open_scratch, st, LC));
        const u32 q0 = 2 * sp.p_lo(g), nqc = 2 * sp.p_hi(g) - q0;
        if (nqc) RC(launch_open(qcoef + (u64)q0 * N, N, nqc, N, w.U1q, w.U1q, oq + q0, w.dummy_open, w.open_scratch, st, LC));
      }
      if (ctx->comm && comm_all_reduce_sum_u32(ctx->comm, w.proof + L.open_t, (2 * WA + QW) * 4, st, &ctx->err) != 0) return ZKIR_ERR_NCCL;
    } else {
      RC(launch_open(w.coef, N, (u32)WA, N, w.U1, w.U2, ot, otg, w.open_scratch, st, LC));
      RC(launch_open(qcoef, N, QW, N, w.U1q, w.U1q, oq, w.dummy_open, w.open_scratch, st, LC));
    }
    RC(launch_observe_hash_tree(w.proof + L.open_t, (u32)(2 * WA + QW) * 4, w.otree, w.chal, c_afri, 4, st, LC));   // observe the tree hash, sample gamma
    DeepArgs da;
    da.lde = w.lde; da.M = M; da.width = (u32)WA; da.qlde = w.qlde; da.qwidth = QW; da.log_n = log_n; da.log_b = p->log_blowup; da.xs = w.xs; da.zeta = c_zeta;
    da.g_mont = bb_to_mont_c(g); da.alpha_fri = c_afri; da.open_t = ot; da.open_tg = otg; da.open_q = oq; da.afp_scratch = w.afp;
    da.out = w.h_layers[0];
    if (sp.on) {
      // row-sharded; FRI layer 0 (natural order, ext4) is completed with one all-gather
      for (u32 g = sp.lo; g < sp.hi; g++) {
        da.seg_log_nj = sp.log_nj; da.seg_j0 = (u64)g * sp.nj;
        RC(launch_deep(da, st, LC));
      }
      if (ctx->comm && comm_all_gather_u32(ctx->comm, reinterpret_cast<u32*>(w.h_layers[0]), 4 * (M / sp.G), st, &ctx->err) != 0) return ZKIR_ERR_NCCL;
    } else {
      RC(launch_deep(da, st, LC));
    }
  }
  // ---- 5. FRI commit phase
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_FRI], st));
  u32 l_sl[32] = {0};
  {
    const u32* inv_w = ntt_powers_table(ctx->tables, hinv(ZKIR_BB_ROOTS[log_m]), 1, M / 2);
    if (!inv_w) { ctx->err = "table alloc"; return ZKIR_ERR_OOM; }
    u32 lshift = shift;
    u32 level = 0;   // fold level: layer `level` has M >> level values
    for (u32 t = 0; t < L.R; t++) {
      const u32 la = t < log_n / 3 ? 3 : log_n % 3;
      const u64 qn = (M >> level) >> la;
      // leaves hash(f[i] || f[i+qn] || ... || f[i+(2^la-1)qn]) + tree + root -> proof, observe, sample beta_t
      if ((crc = commit_tree(ctx, nullptr, 0, 0, reinterpret_cast<const u32*>(w.h_layers[level]), w.h_ltrees[level], qn, w.proof + L.fri_roots + 8 * t,
                             c_betas + 4 * t, 4, &l_sl[t], -1, 1u << la)) != 0) return crc;
      // half-folds with beta, beta^2, beta^4 on the coset, its square, its fourth power: one launch per round (stark.cu), the
      // intermediate layers stay in registers
      u32 cs[3] = {0, 0, 0};
      u32 ls = lshift;
      for (u32 s = 0; s < la; s++) { cs[s] = bb_to_mont_c(hinv(hmul(2, ls))); ls = hmul(ls, ls); }
      RC(launch_fri_fold_multi(w.h_layers[level], w.h_layers[level + la], M >> level, la, c_betas + 4 * t, inv_w, 1u << level, cs, st, LC));
      lshift = ls;
      level += la;
    }
    CU(cudaMemcpyAsync(w.proof + L.final_, w.h_layers[R], 16, cudaMemcpyDeviceToDevice, st));
    RC(launch_challenger(w.chal, w.proof + L.final_, 4, nullptr, 0, 0, st, LC));
    RC(launch_pow_grind(w.chal, p->pow_bits, c_pow_raw, st, LC));
    RC(launch_map(w.proof + L.pow_, c_pow_raw, 1, 1, st, LC));
    RC(launch_challenger(w.chal, w.proof + L.pow_, 1, c_pow_sample, 1, p->pow_bits, st, LC));
    RC(launch_challenger(w.chal, nullptr, 0, w.indices, p->num_queries, log_m, st, LC));
  }
  // ---- 6. queries, canonicalise, D2H
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_QUERIES_D2H], st));
  {
    QueryArgs qa;
    qa.indices = w.indices; qa.num_queries = p->num_queries; qa.log_m = log_m; qa.width = (u32)W; qa.log_n = log_n;
    qa.lde = w.lde; qa.ttree = w.ttree; qa.qlde = w.qlde; qa.qtree = w.qtree;
    qa.aux_width = (u32)AW; qa.atree = w.atree; qa.atree_sl = a_sl; qa.log_lr = L.log_lr;
    qa.layers = w.d_layers; qa.ltrees = w.d_ltrees; qa.fold8_rounds = log_n / 3; qa.last_log_arity = log_n % 3; qa.fri_rounds = (u32)L.R; qa.out = w.proof + L.queries; qa.words_per_query = (u32)L.per_query;
    // sharded trees: the bottom path levels of a leaf exist only on its owner; every rank writes the pieces it owns (zeros
    // elsewhere, rank 0 also everything that is replicated) and one all-reduce assembles the query section ("query gather")
    qa.shard_lo = ctx->shard_lo; qa.shard_hi = ctx->shard_hi; qa.ttree_sl = t_sl; qa.qtree_sl = q_sl; qa.lde_sl = sp.on ? t_sl : 0; qa.qlde_sl = sp.on ? q_sl : 0;
    for (u32 r = 0; r < 32; r++) qa.layer_sl[r] = l_sl[r];
    RC(launch_queries(qa, st, LC));
    if (ctx->comm && L.nq && comm_all_reduce_sum_u32(ctx->comm, w.proof + L.queries, L.per_query * L.nq, st, &ctx->err) != 0) return ZKIR_ERR_NCCL;
    RC(launch_map(w.proof + 8 + np, w.proof + 8 + np, L.total - 8 - np, 0, st, LC));
    CU(cudaMemcpyAsync(w.h_proof, w.proof, L.total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 16, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_COUNT], st));
  return 0;
}

static int expand_error(zkir_ctx* ctx) {  // after the stream is drained: converter error first, then the lookup balance
  for (int k = 0; k < 2; k++) {
    const u64 e = ctx->h_err[k];
    if (e == ~0ull) continue;
    const u32 code = (u32)(e & 0xff);
    ctx->err = "AIR v2 cannot constrain row " + std::to_string((unsigned long long)(e >> 8)) + ": " +
               (code == 8 ? "the lookup fractions do not balance (a chunk outside the range table, an instruction outside the program, or an I/O transcript that is not the trace's)" : pack_err_text(code));
    return ZKIR_ERR_AIR;
  }
  return 0;
}

static int finish_proof(zkir_ctx* ctx, const zkir_params* p, u32 log_n, uint8_t** proof, size_t* proof_len) {
  CU(cudaStreamSynchronize(ctx->stream));
  const Layout L = make_layout(p, log_n);
  uint8_t* out = (uint8_t*)malloc(L.total * 4);
  if (!out) { ctx->err = "malloc"; return ZKIR_ERR_OOM; }
  { int erc = expand_error(ctx); if (erc) { free(out); *proof = nullptr; *proof_len = 0; return erc; } }
  memcpy(out, ctx->ws.h_proof, L.total * 4);
  *proof = out; *proof_len = L.total * 4;
  if (ctx->ws.graph_run) {   // the stage events of a replayed graph are not re-recorded
    ctx->have_stage = false;
  } else {
    for (int s = 0; s < ZKIR_STAGE_COUNT; s++) cudaEventElapsedTime(&ctx->stage_ms[s], ctx->ev[s], ctx->ev[s + 1]);
    ctx->have_stage = true;
  }
  return 0;
}

extern "C" {

int zkir_b200_create(zkir_ctx** out, int device_id) {
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { g_last_error = std::string("no CUDA device: ") + cudaGetErrorString(e); return ZKIR_ERR_CUDA; }
  if (device_id < 0 || device_id >= n) { g_last_error = "device id out of range"; return ZKIR_ERR_ARG; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess || prop.major != 10) {
    g_last_error = "zkir_b200 kernels are built for sm_100a only; device is not compute capability 10.x";
    return ZKIR_ERR_CUDA;
  }
  if (cudaSetDevice(device_id) != cudaSuccess) { g_last_error = "cudaSetDevice failed"; return ZKIR_ERR_CUDA; }
  zkir_ctx* ctx = new zkir_ctx();
  ctx->device = device_id;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; g_last_error = "stream"; return ZKIR_ERR_CUDA; }
  for (int i = 0; i <= ZKIR_STAGE_COUNT; i++) cudaEventCreate(&ctx->ev[i]);
  cudaEventCreate(&ctx->tev[0]); cudaEventCreate(&ctx->tev[1]);
  if (poseidon2_init_constants() != 0) { g_last_error = "constant upload failed"; delete ctx; return ZKIR_ERR_CUDA; }
  if (cudaMalloc(&ctx->d_err, 16) != cudaSuccess || cudaMallocHost(&ctx->h_err, 16) != cudaSuccess) { g_last_error = "error words"; delete ctx; return ZKIR_ERR_OOM; }
  ctx->h_err[0] = ctx->h_err[1] = ~0ull;
  if (cudaMalloc(&ctx->d_io, 64) != cudaSuccess) { g_last_error = "io buffer"; delete ctx; return ZKIR_ERR_OOM; }
  ctx->d_io_cap = 16;
  ctx->tables = ntt_tables_create(ctx->stream, &ctx->launches);
  ctx->fast = fast_ntt_create(ctx->stream, &ctx->launches);
  *out = ctx;
  return 0;
}

void zkir_b200_destroy(zkir_ctx* ctx) {
  if (!ctx) return;
  for (zkir_ctx* w : ctx->workers) zkir_b200_destroy(w);
  ctx->workers.clear();
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx->comm);
  ctx->comm = nullptr;
  ws_free(ctx);
  if (ctx->ntt_tmp) cudaFree(ctx->ntt_tmp);
  for (int i = 0; i < 2; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->rows_dev) cudaFree(ctx->rows_dev);
  if (ctx->full_stage) cudaFreeHost(ctx->full_stage);
  if (ctx->memlog_pinned) cudaFreeHost(ctx->memlog_pinned);
  if (ctx->d_io) cudaFree(ctx->d_io);
  if (ctx->log_pinned) cudaFreeHost(ctx->log_pinned);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
  if (ctx->d_err) cudaFree(ctx->d_err);
  if (ctx->h_err) cudaFreeHost(ctx->h_err);
  ntt_tables_destroy(ctx->tables);
  fast_ntt_destroy(ctx->fast);
  for (int i = 0; i <= ZKIR_STAGE_COUNT; i++) cudaEventDestroy(ctx->ev[i]);
  cudaEventDestroy(ctx->tev[0]); cudaEventDestroy(ctx->tev[1]);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int zkir_b200_set_program(zkir_ctx* ctx, const uint32_t* code, size_t n_code) {
  if (!ctx || (!code && n_code) || n_code > (1ull << 26)) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->program_version && ctx->code.size() == n_code && (n_code == 0 || !memcmp(ctx->code.data(), code, n_code * 4))) return 0;   // unchanged
  ctx->code.assign(code, code + n_code);
  zkir_program_digest(ctx->code.data(), n_code, ctx->code_digest);
  ctx->program_version++;
  for (zkir_ctx* w : ctx->workers) { int rc = zkir_b200_set_program(w, code, n_code); if (rc) return rc; }
  return 0;
}

int zkir_b200_set_io(zkir_ctx* ctx, const uint32_t* events, size_t n_events) {
  if (!ctx || (!events && n_events) || n_events > (1ull << 26)) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->io.assign(events, events + 4 * n_events);
  ctx->io_digest_valid = false;
  if (ctx->d_io_cap < 4 * n_events + 4) {
    if (ctx->d_io) cudaFree(ctx->d_io);
    ctx->d_io = nullptr; ctx->d_io_cap = 0;
    CU(cudaMalloc(&ctx->d_io, (4 * n_events + 4) * 4));
    ctx->d_io_cap = 4 * n_events + 4;
  }
  // ON THE CONTEXT'S STREAM: a plain cudaMemcpy from pageable memory returns once the data is staged, and its DMA is ordered in the legacy
  // default stream, which a non-blocking stream does not wait for -- the next proof could read the previous transcript
  if (n_events) CU(cudaMemcpyAsync(ctx->d_io, ctx->io.data(), 4 * n_events * 4, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

const char* zkir_b200_last_error(const zkir_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_last_error.c_str();
}

void* zkir_b200_alloc_pinned(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr; }
void zkir_b200_free_pinned(void* p) { if (p) cudaFreeHost(p); }
void zkir_b200_free_proof(uint8_t* p) { free(p); }

int zkir_b200_prove(zkir_ctx* ctx, const zkir_params* p, const uint32_t* trace_cols, uint32_t log_n, const uint32_t* pv,
                    uint8_t** proof, size_t* proof_len) {
  if (!ctx) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = check_params(ctx, p, log_n);
  if (rc) return rc;
  if (!trace_cols || !pv || !proof || !proof_len) { ctx->err = "null argument"; return ZKIR_ERR_ARG; }
  if ((rc = ws_prepare(ctx, p, log_n)) != 0) return rc;
  if ((rc = ensure_public_columns(ctx, p, log_n)) != 0) return rc;
  Workspace& w = ctx->ws;
  const size_t trace_bytes = ((size_t)p->width << log_n) * 4;
  w.graph_run = false;
  CU(cudaMemsetAsync(ctx->d_err, 0xff, 16, ctx->stream));
  static const int graph_max_log_n = getenv("ZKIR_GRAPH_MAX_LOG_N") ? atoi(getenv("ZKIR_GRAPH_MAX_LOG_N")) : 12;   // 0 disables
  if ((int)log_n <= graph_max_log_n && !ctx->comm && ctx->shards == 1 && !w.graph_failed && w.proofs_done >= 1) {
    if (!w.h_trace_stage) CU(cudaMallocHost(&w.h_trace_stage, trace_bytes));
    memcpy(w.h_trace_stage, trace_cols, trace_bytes);
    if (w.gexec && (w.graph_pow_bits != p->pow_bits || w.graph_n_io != ctx->io.size() || w.graph_d_io != ctx->d_io)) { cudaGraphExecDestroy(w.gexec); w.gexec = nullptr; }
    if (!w.gexec) {
      // capture once; thread-local mode: other contexts (prove_batch workers) keep allocating and launching meanwhile
      cudaGraph_t graph = nullptr;
      const u64 launches_before = ctx->launches;
      CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      cudaError_t ce = cudaMemcpyAsync(w.trace, w.h_trace_stage, trace_bytes, cudaMemcpyHostToDevice, ctx->stream);
      rc = ce == cudaSuccess ? prove_resident(ctx, p, log_n, pv, w.trace) : ZKIR_ERR_CUDA;
      cudaError_t ee = cudaStreamEndCapture(ctx->stream, &graph);
      w.graph_launches = ctx->launches - launches_before;
      ctx->launches = launches_before;   // nothing ran yet: the replay below counts
      w.graph_pow_bits = p->pow_bits; w.graph_n_io = ctx->io.size(); w.graph_d_io = ctx->d_io;
      if (rc == 0 && ee == cudaSuccess && graph && cudaGraphInstantiate(&w.gexec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
      } else {                     // fall back to ordinary launches for this shape (still the CUDA path, never a CPU path)
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        w.gexec = nullptr; w.graph_failed = true;
        if (rc == ZKIR_ERR_ARG) return rc;
      }
    }
    if (w.gexec) {
      if ((rc = fill_header_stage(ctx, p, log_n, pv)) != 0) return rc;
      CU(cudaGraphLaunch(w.gexec, ctx->stream));
      ctx->launches += w.graph_launches;
      w.graph_run = true;
      w.proofs_done++;
      return finish_proof(ctx, p, log_n, proof, proof_len);
    }
  }
  const u64 l0 = ctx->launches;
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_H2D], ctx->stream));
  CU(cudaMemcpyAsync(w.trace, trace_cols, trace_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = prove_resident(ctx, p, log_n, pv, w.trace)) != 0) return rc;
  w.graph_launches = ctx->launches - l0;
  w.proofs_done++;
  return finish_proof(ctx, p, log_n, proof, proof_len);
}

int zkir_b200_prove_device(zkir_ctx* ctx, const zkir_params* p, const uint32_t* d_trace, uint32_t log_n, const uint32_t* pv,
                           uint8_t** proof, size_t* proof_len) {
  if (!ctx) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = check_params(ctx, p, log_n);
  if (rc) return rc;
  if (!d_trace || !pv || !proof || !proof_len) { ctx->err = "null argument"; return ZKIR_ERR_ARG; }
  if ((rc = ws_prepare(ctx, p, log_n)) != 0) return rc;
  if ((rc = ensure_public_columns(ctx, p, log_n)) != 0) return rc;
  ctx->ws.graph_run = false;   // only zkir_b200_prove replays graphs; a stale flag would hide this proof's stage timings
  CU(cudaMemsetAsync(ctx->d_err, 0xff, 16, ctx->stream));
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_H2D], ctx->stream));
  if ((rc = prove_resident(ctx, p, log_n, pv, d_trace)) != 0) return rc;  // read in place: no pass writes the caller's matrix
  return finish_proof(ctx, p, log_n, proof, proof_len);
}

// {entry_pc, num_cycles, exit_lo, exit_hi, halted}: halted = the run ended in EXIT / EBREAK (not cut by the cycle limit)
static void fill_public_values(u32* pv, u32 entry_point, u64 n_rows, u64 exit_code, int halt_kind) {
  const u64 LIMB = (1u << 20) - 1;
  const u64 code = halt_kind == ZKIR_HALT_EXIT ? exit_code : 0;
  pv[0] = entry_point; pv[1] = (u32)(n_rows % BB_P); pv[2] = (u32)(code & LIMB); pv[3] = (u32)((code >> 20) & LIMB);
  pv[4] = halt_kind == ZKIR_HALT_CYCLE_LIMIT ? 0u : 1u;
}

// H2D of the raw rows + the device converter; columns land in d_cols
static int expand_rows(zkir_ctx* ctx, const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs, uint64_t T, const uint64_t* final_regs,
                       uint64_t final_pc, uint32_t log_n, u32* d_cols, u32 col_lo = 0, u32 col_hi = 0xffffffffu) {
  const u64 N = 1ull << log_n;
  if (T >= N || !pcs || !instrs || !regs || !final_regs) { ctx->err = "bad rows: need n_rows < 2^log_n (the last row is a padding row) and non-null arrays"; return ZKIR_ERR_ARG; }
  const size_t need = T * (8 + 4 + 128) + 64;
  if (ctx->rows_bytes < need) {
    if (ctx->rows_dev) cudaFree(ctx->rows_dev);
    ctx->rows_dev = nullptr; ctx->rows_bytes = 0;
    CU(cudaMalloc(&ctx->rows_dev, need));
    ctx->rows_bytes = need;
  }
  char* base = (char*)ctx->rows_dev;
  u64* d_regs = (u64*)base;
  u64* d_pcs = (u64*)(base + T * 128);
  u32* d_ins = (u32*)(base + T * 136);
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(d_regs, regs, T * 128, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_pcs, pcs, T * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_ins, instrs, T * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(ctx->d_err, 0xff, 16, st));
  ExpandArgs ea;
  ea.pcs = d_pcs; ea.ins = d_ins; ea.regs = d_regs; ea.T = T; ea.N = N;
  for (int k = 0; k < 16; k++) ea.final_regs[k] = final_regs[k];
  ea.final_pc = final_pc; ea.cols = d_cols; ea.err = ctx->d_err; ea.col_lo = col_lo; ea.col_hi = col_hi; ea.n_code = (u32)ctx->code.size();
  RC(launch_trace_expand(ea, st, &ctx->launches));
  return 0;
}
// full profile: rows + host memory replay -> device converter (trace_expand.cu, full build) + boundary cells of the memory argument
extern "C" int zkir_mem_replay_full(const uint32_t*, const uint64_t*, uint64_t, const uint64_t*, const uint32_t*, size_t, uint64_t*, uint32_t*, uint64_t*, uint64_t*);
extern "C" int zkir_mem_boundary_full(uint32_t, uint32_t*, uint64_t, uint32_t*, uint32_t*);
static const u32 MEM_BOUNDARY_COL0 = ZKIR_PROFILE_FULL_COL_BOUNDARY0, MEM_BOUNDARY_COLS = ZKIR_PROFILE_FULL_BOUNDARY_COLS;   // img_fin0 .. ram_fin_ts of the full table
// boundary cells of the memory argument, after the converter kernel (which zeroed those columns): the first n_img rows of the image
// columns, the first n_ram rows of the RAM columns, and the boundary's own lookup multiplicities (d_delta: [1024] range + [128] 7-bit)
static int upload_mem_boundary(zkir_ctx* ctx, const u32* h_bcols, u64 bstride, u64 n_img, u64 n_ram, const u32* d_delta, u32* d_cols, u64 N) {
  cudaStream_t st = ctx->stream;
  for (u32 c = 0; c < MEM_BOUNDARY_COLS; c++) {
    const u64 n = c < 9 ? n_img : n_ram;   // img_fin0..7, img_fin_ts (9 columns) | ram_on, ram_a0..2, ram_e0..2, ram_fin0..7, ram_fin_ts (16)
    if (n) CU(cudaMemcpyAsync(d_cols + (u64)(MEM_BOUNDARY_COL0 + c) * N, h_bcols + (size_t)c * bstride, n * 4, cudaMemcpyHostToDevice, st));
  }
  RC(launch_add_u32(d_cols + (u64)ZKIR_COL_M_RNG * N, d_delta, 1024, st, &ctx->launches));
  RC(launch_add_u32(d_cols + (u64)ZKIR_PROFILE_FULL_COL_M_B7 * N, d_delta + 1024, 128, st, &ctx->launches));
  return 0;
}
static int expand_rows_full(zkir_ctx* ctx, const uint32_t* instrs, const uint64_t* pcs, const uint64_t* regs, uint64_t T, const uint64_t* final_regs,
                            uint64_t final_pc, uint32_t log_n, u32* d_cols) {
  const u64 N = 1ull << log_n;
  if (T >= N || !pcs || !instrs || !regs || !final_regs) { ctx->err = "bad rows: need n_rows < 2^log_n (the last row is a padding row) and non-null arrays"; return ZKIR_ERR_ARG; }
  // pinned staging: old_word[T] | prev_ts[T] | multiplicity deltas [1024 + 128] | boundary columns [25][bstride]
  auto stage = [&](size_t bytes) -> int {
    if (ctx->full_stage_bytes < bytes) {
      if (ctx->full_stage) cudaFreeHost(ctx->full_stage);
      ctx->full_stage = nullptr; ctx->full_stage_bytes = 0;
      CU(cudaMallocHost(&ctx->full_stage, bytes));
      ctx->full_stage_bytes = bytes;
    }
    return 0;
  };
  const size_t fixed = T * 12 + 64 + (1024 + 128) * 4;
  RC(stage(fixed + (size_t)MEM_BOUNDARY_COLS * 4096 * 4));
  char* hb = (char*)ctx->full_stage;
  u64* h_old = (u64*)hb;
  u32* h_pts = (u32*)(hb + T * 8);
  memset(hb, 0, T * 12);
  u64 n_img = 0, n_ram = 0;
  int rc = zkir_mem_replay_full(instrs, regs, T, final_regs, ctx->code.data(), ctx->code.size(), h_old, h_pts, &n_img, &n_ram);
  if (rc) { ctx->err = zkir_b200_last_error(nullptr); return rc; }
  const u64 bstride = std::max<u64>(std::max(n_img, n_ram), 1);
  if (bstride > 4096) {   // grow the staging, keep what the replay wrote
    std::vector<char> keep(hb, hb + T * 12);
    RC(stage(fixed + (size_t)MEM_BOUNDARY_COLS * bstride * 4));
    hb = (char*)ctx->full_stage; h_old = (u64*)hb; h_pts = (u32*)(hb + T * 8);
    memcpy(hb, keep.data(), keep.size());
  }
  u32* h_delta = (u32*)(hb + ((T * 12 + 63) & ~(size_t)63));
  u32* h_bcols = h_delta + 1024 + 128;
  memset(h_bcols, 0, (size_t)MEM_BOUNDARY_COLS * bstride * 4);
  rc = zkir_mem_boundary_full(log_n, h_bcols, bstride, h_delta, h_delta + 1024);
  if (rc) { ctx->err = zkir_b200_last_error(nullptr); return rc; }
  // device staging: regs | pcs | ins | old_word | prev_ts | deltas
  const size_t need = T * (128 + 8 + 4 + 8 + 4) + 64 + (1024 + 128) * 4 + 64;
  if (ctx->rows_bytes < need) {
    if (ctx->rows_dev) cudaFree(ctx->rows_dev);
    ctx->rows_dev = nullptr; ctx->rows_bytes = 0;
    CU(cudaMalloc(&ctx->rows_dev, need));
    ctx->rows_bytes = need;
  }
  char* base = (char*)ctx->rows_dev;
  u64* d_regs = (u64*)base;
  u64* d_pcs = (u64*)(base + T * 128);
  u64* d_old = (u64*)(base + T * 136);
  u32* d_ins = (u32*)(base + T * 144);
  u32* d_pts = (u32*)(base + T * 148);
  u32* d_delta = (u32*)(base + ((T * 152 + 63) & ~(size_t)63));
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(d_regs, regs, T * 128, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_pcs, pcs, T * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_ins, instrs, T * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_old, h_old, T * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_pts, h_pts, T * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_delta, h_delta, (1024 + 128) * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(ctx->d_err, 0xff, 16, st));
  ExpandFullArgs fa;
  fa.rows.pcs = d_pcs; fa.rows.ins = d_ins; fa.rows.regs = d_regs; fa.rows.T = T; fa.rows.N = N;
  for (int k = 0; k < 16; k++) fa.rows.final_regs[k] = final_regs[k];
  fa.rows.final_pc = final_pc; fa.rows.cols = d_cols; fa.rows.err = ctx->d_err; fa.rows.n_code = (u32)ctx->code.size();
  fa.old_word = d_old; fa.prev_ts = d_pts;
  RC(launch_trace_expand_full(fa, st, &ctx->launches));
  return upload_mem_boundary(ctx, h_bcols, bstride, n_img, n_ram, d_delta, d_cols, N);
}
int zkir_b200_prove_rows(zkir_ctx* ctx, const zkir_params* p, const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs,
                         uint64_t n_rows, const uint64_t* final_regs, uint64_t final_pc, uint32_t entry_point, uint64_t exit_code,
                         int halt_kind, uint32_t log_n, uint32_t* pv_out, uint8_t** proof, size_t* proof_len) {
  if (!ctx) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = check_params(ctx, p, log_n);
  if (rc) return rc;
  if (!pv_out || !proof || !proof_len) { ctx->err = "null argument"; return ZKIR_ERR_ARG; }
  if ((rc = ws_prepare(ctx, p, log_n)) != 0) return rc;
  if ((rc = ensure_public_columns(ctx, p, log_n)) != 0) return rc;
  ctx->ws.graph_run = false;   // only zkir_b200_prove replays graphs; a stale flag would hide this proof's stage timings
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_H2D], ctx->stream));
  if (profile_is_full(p->width)) {
    // full profile: the host only replays the run's memory in program order (the one sequential step: which word and which previous
    // timestamp every load / store sees); the 248-column table is expanded on the device from the rows, like the core converter
    if ((rc = expand_rows_full(ctx, instrs, pcs, regs, n_rows, final_regs, final_pc, log_n, ctx->ws.trace)) != 0) return rc;
    fill_public_values(pv_out, entry_point, n_rows, exit_code, halt_kind);
    if ((rc = prove_resident(ctx, p, log_n, pv_out, ctx->ws.trace)) != 0) return rc;
    return finish_proof(ctx, p, log_n, proof, proof_len);
  }
  u32 c_lo, c_hi;
  trace_col_range(ctx, p, log_n, &c_lo, &c_hi);   // a sharded proof only materialises the columns this rank transforms
  if ((rc = expand_rows(ctx, pcs, instrs, regs, n_rows, final_regs, final_pc, log_n, ctx->ws.trace, c_lo, c_hi)) != 0) return rc;
  fill_public_values(pv_out, entry_point, n_rows, exit_code, halt_kind);
  if ((rc = prove_resident(ctx, p, log_n, pv_out, ctx->ws.trace)) != 0) return rc;
  return finish_proof(ctx, p, log_n, proof, proof_len);
}

// write-log flavour: H2D of (pc32, word, wlog) + last-writer scan + converter.  Device staging: wlog[Tc] | pcs[Tc] | ins[Tc] | scan scratch
struct WlStage { u64* d_wlog; u32* d_pcs; u32* d_ins; int* d_scr; };
static int wl_stage(zkir_ctx* ctx, u64 Tc, u64 n_scan_rows, WlStage* o) {
  const size_t scratch = trace_expand_wl_scratch_ints(n_scan_rows) * 4;
  const size_t need = Tc * 16 + scratch + 64;
  if (ctx->rows_bytes < need) {
    if (ctx->rows_dev) cudaFree(ctx->rows_dev);
    ctx->rows_dev = nullptr; ctx->rows_bytes = 0;
    CU(cudaMalloc(&ctx->rows_dev, need));
    ctx->rows_bytes = need;
  }
  char* base = (char*)ctx->rows_dev;
  o->d_wlog = (u64*)base;
  o->d_pcs = (u32*)(base + Tc * 8);
  o->d_ins = (u32*)(base + Tc * 12);
  o->d_scr = (int*)(base + ((Tc * 16 + 15) & ~(size_t)15));
  return 0;
}
static int wl_convert(zkir_ctx* ctx, const WlStage& sg, uint64_t T, uint64_t final_pc, uint32_t log_n, u32* d_cols, u32 col_lo, u32 col_hi) {
  cudaStream_t st = ctx->stream;
  CU(cudaMemsetAsync(ctx->d_err, 0xff, 16, st));
  WlArgs wa;
  wa.pcs = sg.d_pcs; wa.ins = sg.d_ins; wa.wlog = sg.d_wlog; wa.T = T; wa.N = 1ull << log_n; wa.final_pc = final_pc; wa.chunk_prev = sg.d_scr; wa.cols = d_cols;
  wa.err = ctx->d_err; wa.col_lo = col_lo; wa.col_hi = col_hi; wa.n_code = (u32)ctx->code.size();
  RC(launch_trace_expand_wl(wa, st, &ctx->launches));
  return 0;
}
static int expand_writelog(zkir_ctx* ctx, const uint32_t* pcs, const uint32_t* instrs, const uint64_t* wlog, uint64_t T, uint64_t final_pc,
                           uint32_t log_n, u32* d_cols, u32 col_lo = 0, u32 col_hi = 0xffffffffu) {
  const u64 N = 1ull << log_n;
  if (T >= N || !pcs || !instrs || !wlog) { ctx->err = "bad write log: need n_rows < 2^log_n (the last row is a padding row) and non-null arrays"; return ZKIR_ERR_ARG; }
  // Sharded proof (collective call, every rank holds the same log in host memory): each rank uploads only ITS row segment over
  // PCIe and the segments are exchanged over NVLink with one grouped all-gather, so the host link carries T/G rows per GPU.
  const u32 G = ctx->comm ? ctx->shards : 1;
  const bool split = G > 1 && T >= 65536;
  const u64 chunk = split ? (((T + G - 1) / G + 3) & ~3ull) : T;   // rows per rank
  const u64 Tc = split ? chunk * G : T;                             // capacity of the device arrays
  WlStage sg;
  { int rc = wl_stage(ctx, Tc, N, &sg); if (rc) return rc; }
  cudaStream_t st = ctx->stream;
  if (split) {
    const u64 r0 = (u64)comm_rank(ctx->comm) * chunk, r1 = r0 + chunk < T ? r0 + chunk : T;
    if (r1 > r0) {
      CU(cudaMemcpyAsync(sg.d_wlog + r0, wlog + r0, (r1 - r0) * 8, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(sg.d_pcs + r0, pcs + r0, (r1 - r0) * 4, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(sg.d_ins + r0, instrs + r0, (r1 - r0) * 4, cudaMemcpyHostToDevice, st));
    }
    unsigned* bufs[3] = {reinterpret_cast<unsigned*>(sg.d_wlog), sg.d_pcs, sg.d_ins};
    const size_t per[3] = {(size_t)chunk * 2, (size_t)chunk, (size_t)chunk};
    if (comm_all_gather_group_u32(ctx->comm, bufs, per, 3, st, &ctx->err) != 0) return ZKIR_ERR_NCCL;
  } else {
    CU(cudaMemcpyAsync(sg.d_wlog, wlog, T * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sg.d_pcs, pcs, T * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sg.d_ins, instrs, T * 4, cudaMemcpyHostToDevice, st));
  }
  return wl_convert(ctx, sg, T, final_pc, log_n, d_cols, col_lo, col_hi);
}

int zkir_b200_prove_writelog(zkir_ctx* ctx, const zkir_params* p, const uint32_t* pcs, const uint32_t* instrs, const uint64_t* wlog,
                             uint64_t n_rows, uint64_t final_pc, uint32_t entry_point, uint64_t exit_code, int halt_kind, uint32_t log_n,
                             uint32_t* pv_out, uint8_t** proof, size_t* proof_len) {
  if (!ctx) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = check_params(ctx, p, log_n);
  if (rc) return rc;
  if (!pv_out || !proof || !proof_len) { ctx->err = "null argument"; return ZKIR_ERR_ARG; }
  if (profile_is_full(p->width)) { ctx->err = "the register write log alone does not describe loads and stores: full-profile runs take zkir_b200_prove_writelog_mem (write log + memory log), zkir_b200_prove_rows or zkir_b200_prove_program"; return ZKIR_ERR_ARG; }
  if ((rc = ws_prepare(ctx, p, log_n)) != 0) return rc;
  if ((rc = ensure_public_columns(ctx, p, log_n)) != 0) return rc;
  ctx->ws.graph_run = false;   // only zkir_b200_prove replays graphs; a stale flag would hide this proof's stage timings
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_H2D], ctx->stream));
  u32 c_lo, c_hi;
  trace_col_range(ctx, p, log_n, &c_lo, &c_hi);   // a sharded proof only materialises the columns this rank transforms
  if ((rc = expand_writelog(ctx, pcs, instrs, wlog, n_rows, final_pc, log_n, ctx->ws.trace, c_lo, c_hi)) != 0) return rc;
  fill_public_values(pv_out, entry_point, n_rows, exit_code, halt_kind);
  if ((rc = prove_resident(ctx, p, log_n, pv_out, ctx->ws.trace)) != 0) return rc;
  return finish_proof(ctx, p, log_n, proof, proof_len);
}

// Program -> Proof.  The interpreter (host/vm.cc) runs in the calling thread and appends the write log to pinned memory; every
// PROG_CHUNK cycles the finished rows start their host->device copy on a second stream, so at the end of the run only the last
// chunk is left to move.  Then: last-writer scan + converter + proof, as in zkir_b200_prove_writelog.
#define PROG_CHUNK (1u << 16)
extern "C" int zkir_vm_run_writelog_cb(const uint32_t*, size_t, const uint8_t*, size_t, uint32_t, const uint64_t*, size_t, uint64_t, uint32_t*, uint32_t*,
                                       uint64_t*, uint64_t, void (*)(void*, uint64_t), void*, uint64_t, zkir_vm_result**);
struct ProgUpload {
  zkir_ctx* ctx; const u32 *h_pcs, *h_ins; const u64* h_wlog; WlStage sg; u64 done; cudaError_t e;
  const u64* h_old = nullptr; const u32* h_pts = nullptr; u64* d_old = nullptr; u32* d_pts = nullptr;   // memory log (full profile)
};
static void prog_upload_to(ProgUpload* u, u64 rows_done) {
  if (rows_done <= u->done || u->e != cudaSuccess) return;
  const u64 r0 = u->done, n = rows_done - r0;
  cudaStream_t cs = u->ctx->copy_stream;
  cudaError_t e = cudaMemcpyAsync(u->sg.d_wlog + r0, u->h_wlog + r0, n * 8, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess) e = cudaMemcpyAsync(u->sg.d_pcs + r0, u->h_pcs + r0, n * 4, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess) e = cudaMemcpyAsync(u->sg.d_ins + r0, u->h_ins + r0, n * 4, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess && u->h_old) e = cudaMemcpyAsync(u->d_old + r0, u->h_old + r0, n * 8, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess && u->h_old) e = cudaMemcpyAsync(u->d_pts + r0, u->h_pts + r0, n * 4, cudaMemcpyHostToDevice, cs);
  u->e = e; u->done = rows_done;
}
static void prog_on_chunk(void* user, uint64_t rows_done) { prog_upload_to(static_cast<ProgUpload*>(user), rows_done); }

// Full profile, Program -> Proof: the interpreter appends the register write log AND the memory log (the word before each load / store
// and that word's previous timestamp: 16 + 12 B per cycle) to pinned memory, chunks upload while it runs, the device rebuilds the
// registers and expands the 248-column table; the boundary cells come from the interpreter's list of touched words.  No host replay.
extern "C" int zkir_vm_run_writelog_mem_cb(const uint32_t*, size_t, const uint8_t*, size_t, uint32_t, const uint64_t*, size_t, uint64_t, uint32_t*, uint32_t*,
                                           uint64_t*, uint64_t*, uint32_t*, uint64_t, void (*)(void*, uint64_t), void*, uint64_t, zkir_vm_result**);
extern "C" int zkir_mem_boundary_from_words_full(const uint32_t*, size_t, const uint64_t*, const uint64_t*, const uint32_t*, size_t, uint32_t, uint32_t*, uint64_t,
                                                 uint32_t*, uint32_t*);
// device staging of the full profile's logs: wlog | old | pcs | ins | pts [cap each] | scan scratch | multiplicity deltas
struct FullWlStage { WlStage sg; u64* d_old; u32* d_pts; u32* d_delta; };
static int full_wl_stage(zkir_ctx* ctx, u64 cap, u64 n_scan_min, FullWlStage* fs) {
  u64 n_scan = 1ull << ZKIR_RANGE_BITS;
  while (n_scan < cap || n_scan < n_scan_min) n_scan <<= 1;
  const size_t scratch = trace_expand_wl_scratch_ints(n_scan) * 4;
  const size_t need = cap * 28 + scratch + 64 + (1024 + 128) * 4 + 64;
  if (ctx->rows_bytes < need) {
    if (ctx->rows_dev) cudaFree(ctx->rows_dev);
    ctx->rows_dev = nullptr; ctx->rows_bytes = 0;
    CU(cudaMalloc(&ctx->rows_dev, need));
    ctx->rows_bytes = need;
  }
  char* db = (char*)ctx->rows_dev;
  fs->sg.d_wlog = (u64*)db; fs->d_old = (u64*)(db + cap * 8);
  fs->sg.d_pcs = (u32*)(db + cap * 16); fs->sg.d_ins = (u32*)(db + cap * 20); fs->d_pts = (u32*)(db + cap * 24);
  fs->sg.d_scr = (int*)(db + ((cap * 28 + 15) & ~(size_t)15));
  fs->d_delta = (u32*)(db + ((cap * 28 + scratch + 63) & ~(size_t)63));
  return 0;
}
// logs on the device (copies ordered before ctx->stream's next work) -> proof: boundary cells from the touched words (host, a few rows),
// register rebuild + 248-column converter, boundary upload, proof.  ws_prepare / check_params done by the caller.
static int prove_full_from_logs(zkir_ctx* ctx, const zkir_params* p, const FullWlStage& fs, u64 T, u64 final_pc, const uint64_t* widx, const uint64_t* word,
                                const uint32_t* ts, size_t n_words, uint32_t log_n, uint32_t entry_point, uint64_t exit_code, int halt_kind, uint32_t* pv_out,
                                uint8_t** proof, size_t* proof_len) {
  int rc;
  const u64 N = 1ull << log_n;
  const size_t n_code = ctx->code.size();
  const u64 n_img = zkir_image_words(n_code);
  u64 n_ram = 0;
  for (size_t k = 0; k < n_words; k++) n_ram += widx[k] >= n_img;
  const u64 bstride = std::max<u64>(std::max(n_img, n_ram), 1);
  const size_t bbytes = (1024 + 128) * 4 + (size_t)MEM_BOUNDARY_COLS * bstride * 4;
  CU(cudaStreamSynchronize(ctx->stream));   // a previous proof's boundary upload may still read the pinned staging
  if (ctx->full_stage_bytes < bbytes) {
    if (ctx->full_stage) cudaFreeHost(ctx->full_stage);
    ctx->full_stage = nullptr; ctx->full_stage_bytes = 0;
    cudaError_t me = cudaMallocHost(&ctx->full_stage, bbytes);
    if (me != cudaSuccess) { ctx->err = std::string("cudaMallocHost: ") + cudaGetErrorString(me); return ZKIR_ERR_OOM; }
    ctx->full_stage_bytes = bbytes;
  }
  u32* h_delta = ctx->full_stage;
  u32* h_bcols = h_delta + 1024 + 128;
  memset(h_bcols, 0, (size_t)MEM_BOUNDARY_COLS * bstride * 4);
  rc = zkir_mem_boundary_from_words_full(ctx->code.data(), n_code, widx, word, ts, n_words, log_n, h_bcols, bstride, h_delta, h_delta + 1024);
  if (rc) { ctx->err = zkir_b200_last_error(nullptr); return rc; }
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(fs.d_delta, h_delta, (1024 + 128) * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(ctx->d_err, 0xff, 16, st));
  WlFullArgs wa;
  wa.pcs = fs.sg.d_pcs; wa.ins = fs.sg.d_ins; wa.wlog = fs.sg.d_wlog; wa.old_word = fs.d_old; wa.prev_ts = fs.d_pts; wa.T = T; wa.N = N; wa.final_pc = final_pc;
  wa.chunk_prev = fs.sg.d_scr; wa.cols = ctx->ws.trace; wa.err = ctx->d_err; wa.n_code = (u32)n_code;
  RC(launch_trace_expand_wl_full(wa, st, &ctx->launches));
  if ((rc = upload_mem_boundary(ctx, h_bcols, bstride, n_img, n_ram, fs.d_delta, ctx->ws.trace, N)) != 0) return rc;
  fill_public_values(pv_out, entry_point, T, exit_code, halt_kind);
  if ((rc = prove_resident(ctx, p, log_n, pv_out, ctx->ws.trace)) != 0) return rc;
  return finish_proof(ctx, p, log_n, proof, proof_len);
}

int zkir_b200_prove_writelog_mem(zkir_ctx* ctx, const zkir_params* p, const uint32_t* pcs, const uint32_t* instrs, const uint64_t* wlog,
                                 const uint64_t* mem_old, const uint32_t* mem_pts, uint64_t n_rows, const uint64_t* mem_widx, const uint64_t* mem_word,
                                 const uint32_t* mem_ts, size_t n_words, uint64_t final_pc, uint32_t entry_point, uint64_t exit_code, int halt_kind,
                                 uint32_t log_n, uint32_t* pv_out, uint8_t** proof, size_t* proof_len) {
  if (!ctx) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = check_params(ctx, p, log_n);
  if (rc) return rc;
  if (!profile_is_full(p->width)) { ctx->err = "zkir_b200_prove_writelog_mem is the full profile's entry (params.width = ZKIR_AIR_FULL_WIDTH); the core profile takes zkir_b200_prove_writelog"; return ZKIR_ERR_ARG; }
  if (ctx->comm) { ctx->err = "a sharded context proves full-profile runs from rows (zkir_b200_prove_rows)"; return ZKIR_ERR_ARG; }
  if (!pv_out || !proof || !proof_len || !pcs || !instrs || !wlog || !mem_old || !mem_pts || (n_words && (!mem_widx || !mem_word || !mem_ts))) { ctx->err = "null argument"; return ZKIR_ERR_ARG; }
  if (n_rows >= (1ull << log_n)) { ctx->err = "bad write log: need n_rows < 2^log_n (the last row is a padding row)"; return ZKIR_ERR_ARG; }
  for (size_t k = 1; k < n_words; k++) if (mem_widx[k] <= mem_widx[k - 1]) { ctx->err = "memory log: touched words must be strictly ascending"; return ZKIR_ERR_ARG; }
  if ((rc = ws_prepare(ctx, p, log_n)) != 0) return rc;
  if ((rc = ensure_public_columns(ctx, p, log_n)) != 0) return rc;
  ctx->ws.graph_run = false;
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_H2D], ctx->stream));
  FullWlStage fs;
  if ((rc = full_wl_stage(ctx, std::max<u64>(n_rows, 1), ctx->code.size(), &fs)) != 0) return rc;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(fs.sg.d_wlog, wlog, n_rows * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(fs.d_old, mem_old, n_rows * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(fs.sg.d_pcs, pcs, n_rows * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(fs.sg.d_ins, instrs, n_rows * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(fs.d_pts, mem_pts, n_rows * 4, cudaMemcpyHostToDevice, st));
  return prove_full_from_logs(ctx, p, fs, n_rows, final_pc, mem_widx, mem_word, mem_ts, n_words, log_n, entry_point, exit_code, halt_kind, pv_out, proof, proof_len);
}

static int prove_program_full_wl(zkir_ctx* ctx, const zkir_params* p, const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data,
                                 uint32_t entry_point, const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles, uint32_t* pv_out,
                                 uint64_t* out_cycles, uint32_t* out_log_n, uint8_t** proof, size_t* proof_len) {
  int rc;
  if (ctx->memlog_capacity < max_cycles) {   // 28 B per cycle: wlog, old word | pcs, ins, prev_ts
    if (ctx->memlog_pinned) cudaFreeHost(ctx->memlog_pinned);
    ctx->memlog_pinned = nullptr; ctx->memlog_capacity = 0;
    CU(cudaMallocHost(&ctx->memlog_pinned, max_cycles * 28));
    ctx->memlog_capacity = max_cycles;
  }
  if (!ctx->copy_stream) { CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming)); }
  const u64 cap = ctx->memlog_capacity;
  char* hb = (char*)ctx->memlog_pinned;
  u64* h_wlog = (u64*)hb; u64* h_old = (u64*)(hb + cap * 8);
  u32* h_pcs = (u32*)(hb + cap * 16); u32* h_ins = (u32*)(hb + cap * 20); u32* h_pts = (u32*)(hb + cap * 24);
  CU(cudaStreamSynchronize(ctx->stream));   // the previous proof may still read the device staging
  FullWlStage fs;
  if ((rc = full_wl_stage(ctx, max_cycles, n_code, &fs)) != 0) return rc;
  ProgUpload up;
  up.ctx = ctx; up.h_pcs = h_pcs; up.h_ins = h_ins; up.h_wlog = h_wlog; up.done = 0; up.e = cudaSuccess;
  up.sg = fs.sg; up.d_old = fs.d_old; up.d_pts = fs.d_pts; up.h_old = h_old; up.h_pts = h_pts;
  zkir_vm_result* res = nullptr;
  rc = zkir_vm_run_writelog_mem_cb(code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles, h_pcs, h_ins, h_wlog, h_old, h_pts, cap,
                                   prog_on_chunk, &up, PROG_CHUNK, &res);
  if (rc) { ctx->err = std::string("interpreter: ") + zkir_vm_last_error(); return rc; }
  struct Free { zkir_vm_result* r; ~Free() { zkir_vm_free(r); } } free_res{res};   // the touched-word list is read until the converter is launched
  const u64 T = zkir_vm_cycles(res), final_pc = zkir_vm_final_pc(res);
  const int halt_kind = zkir_vm_halt_kind(res);
  const u64 exit_code = zkir_vm_exit_code(res);
  u32 log_n = ZKIR_RANGE_BITS;
  while ((1ull << log_n) <= T || (1ull << log_n) < n_code) log_n++;   // at least one padding row after the last cycle
  if (out_cycles) *out_cycles = T;
  if (out_log_n) *out_log_n = log_n;
  if ((rc = zkir_b200_set_io(ctx, zkir_vm_io(res), zkir_vm_io_len(res))) != 0) return rc;   // the run's public I/O transcript is part of the statement
  if ((rc = check_params(ctx, p, log_n)) != 0) return rc;
  if ((rc = ws_prepare(ctx, p, log_n)) != 0) return rc;
  if ((rc = ensure_public_columns(ctx, p, log_n)) != 0) return rc;
  ctx->ws.graph_run = false;
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_H2D], ctx->stream));
  prog_upload_to(&up, T);                                   // the tail that no chunk boundary covered
  if (up.e != cudaSuccess) { ctx->err = std::string("write-log upload: ") + cudaGetErrorString(up.e); return ZKIR_ERR_CUDA; }
  CU(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
  CU(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
  return prove_full_from_logs(ctx, p, fs, T, final_pc, zkir_vm_memlog_widx(res), zkir_vm_memlog_word(res), zkir_vm_memlog_ts(res), zkir_vm_memlog_count(res), log_n,
                              entry_point, exit_code, halt_kind, pv_out, proof, proof_len);
}

int zkir_b200_prove_program(zkir_ctx* ctx, const zkir_params* p, const uint32_t* code, size_t n_code, const uint8_t* data, size_t n_data,
                            uint32_t entry_point, const uint64_t* inputs, size_t n_inputs, uint64_t max_cycles,
                            uint32_t* pv_out, uint64_t* out_cycles, uint32_t* out_log_n, uint8_t** proof, size_t* proof_len) {
  if (!ctx || !p || !code || !pv_out || !proof || !proof_len || max_cycles == 0 || max_cycles > (1ull << 26)) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = zkir_b200_set_program(ctx, code, n_code);
  if (rc) return rc;
  if (profile_is_full(p->width) && ctx->comm == nullptr) return prove_program_full_wl(ctx, p, code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles,
                                                                                       pv_out, out_cycles, out_log_n, proof, proof_len);
  if (profile_is_full(p->width)) {
    // full profile on a sharded context: the interpreter records full rows (zkir_vm_run with the execution trace on); zkir_b200_prove_rows
    // replays the run's memory on the host and expands the wide table on the device
    zkir_vm_result* res = nullptr;
    rc = zkir_vm_run(code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles, 1, &res);
    if (rc) { ctx->err = std::string("interpreter: ") + zkir_vm_last_error(); return rc; }
    const u64 T = zkir_vm_cycles(res);
    u32 log_n = ZKIR_RANGE_BITS;
    while ((1ull << log_n) <= T || (1ull << log_n) < n_code) log_n++;
    if (out_cycles) *out_cycles = T;
    if (out_log_n) *out_log_n = log_n;
    rc = zkir_b200_set_io(ctx, zkir_vm_io(res), zkir_vm_io_len(res));
    if (!rc) rc = zkir_b200_prove_rows(ctx, p, zkir_vm_trace_pc(res), zkir_vm_trace_instr(res), zkir_vm_trace_regs(res), T, zkir_vm_final_regs(res),
                                       zkir_vm_final_pc(res), entry_point, zkir_vm_exit_code(res), zkir_vm_halt_kind(res), log_n, pv_out, proof, proof_len);
    zkir_vm_free(res);
    return rc;
  }
  if (ctx->log_capacity < max_cycles) {
    if (ctx->log_pinned) cudaFreeHost(ctx->log_pinned);
    ctx->log_pinned = nullptr; ctx->log_capacity = 0;
    CU(cudaMallocHost(&ctx->log_pinned, max_cycles * 16));
    ctx->log_capacity = max_cycles;
  }
  if (!ctx->copy_stream) { CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming)); }
  u64* h_wlog = (u64*)ctx->log_pinned;
  u32* h_pcs = (u32*)((char*)ctx->log_pinned + ctx->log_capacity * 8);
  u32* h_ins = (u32*)((char*)ctx->log_pinned + ctx->log_capacity * 12);
  CU(cudaStreamSynchronize(ctx->stream));   // the previous proof may still read the device staging
  ProgUpload up;
  up.ctx = ctx; up.h_pcs = h_pcs; up.h_ins = h_ins; up.h_wlog = h_wlog; up.done = 0; up.e = cudaSuccess;
  const bool overlap = ctx->comm == nullptr;   // a sharded proof uploads by row segment inside prove_writelog instead
  // the scan scratch depends on the (still unknown) padded trace length: size it for the largest trace the log can hold
  u64 n_scan = 1ull << ZKIR_RANGE_BITS;
  while (n_scan < max_cycles || n_scan < n_code) n_scan <<= 1;
  if (overlap && (rc = wl_stage(ctx, max_cycles, n_scan, &up.sg)) != 0) return rc;
  zkir_vm_result* res = nullptr;
  rc = zkir_vm_run_writelog_cb(code, n_code, data, n_data, entry_point, inputs, n_inputs, max_cycles, h_pcs, h_ins, h_wlog, ctx->log_capacity,
                               overlap ? prog_on_chunk : nullptr, &up, PROG_CHUNK, &res);
  if (rc) { ctx->err = std::string("interpreter: ") + zkir_vm_last_error(); return rc; }
  const u64 T = zkir_vm_cycles(res), final_pc = zkir_vm_final_pc(res);
  const int halt_kind = zkir_vm_halt_kind(res);
  const u64 exit_code = zkir_vm_exit_code(res);
  rc = zkir_b200_set_io(ctx, zkir_vm_io(res), zkir_vm_io_len(res));   // the run's public I/O transcript is part of the statement
  zkir_vm_free(res);
  if (rc) return rc;
  u32 log_n = ZKIR_RANGE_BITS;
  while ((1ull << log_n) <= T || (1ull << log_n) < n_code) log_n++;   // at least one padding row after the last cycle
  if (out_cycles) *out_cycles = T;
  if (out_log_n) *out_log_n = log_n;
  if (!overlap) return zkir_b200_prove_writelog(ctx, p, h_pcs, h_ins, h_wlog, T, final_pc, entry_point, exit_code, halt_kind, log_n, pv_out, proof, proof_len);
  if ((rc = check_params(ctx, p, log_n)) != 0) return rc;
  if ((rc = ws_prepare(ctx, p, log_n)) != 0) return rc;
  if ((rc = ensure_public_columns(ctx, p, log_n)) != 0) return rc;
  ctx->ws.graph_run = false;
  CU(cudaEventRecord(ctx->ev[ZKIR_STAGE_H2D], ctx->stream));
  prog_upload_to(&up, T);                                   // the tail that no chunk boundary covered
  if (up.e != cudaSuccess) { ctx->err = std::string("write-log upload: ") + cudaGetErrorString(up.e); return ZKIR_ERR_CUDA; }
  CU(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
  CU(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
  if ((rc = wl_convert(ctx, up.sg, T, final_pc, log_n, ctx->ws.trace, 0, 0xffffffffu)) != 0) return rc;
  fill_public_values(pv_out, entry_point, T, exit_code, halt_kind);
  if ((rc = prove_resident(ctx, p, log_n, pv_out, ctx->ws.trace)) != 0) return rc;
  return finish_proof(ctx, p, log_n, proof, proof_len);
}

int zkir_b200_expand_writelog(zkir_ctx* ctx, const uint32_t* pcs, const uint32_t* instrs, const uint64_t* wlog, uint64_t n_rows,
                              uint64_t final_pc, uint32_t log_n, uint32_t* d_cols) {
  if (!ctx || !d_cols || log_n < ZKIR_RANGE_BITS || log_n > 26) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = expand_writelog(ctx, pcs, instrs, wlog, n_rows, final_pc, log_n, d_cols);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return expand_error(ctx);
}

int zkir_b200_expand_rows(zkir_ctx* ctx, const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs, uint64_t n_rows,
                          const uint64_t* final_regs, uint64_t final_pc, uint32_t log_n, uint32_t* d_cols) {
  if (!ctx || !d_cols || log_n < ZKIR_RANGE_BITS || log_n > 26) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = expand_rows(ctx, pcs, instrs, regs, n_rows, final_regs, final_pc, log_n, d_cols);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return expand_error(ctx);
}

int zkir_b200_expand_rows_full(zkir_ctx* ctx, const uint64_t* pcs, const uint32_t* instrs, const uint64_t* regs, uint64_t n_rows,
                               const uint64_t* final_regs, uint64_t final_pc, uint32_t log_n, uint32_t* d_cols) {
  if (!ctx || !d_cols || log_n < ZKIR_RANGE_BITS || log_n > 26) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  int rc = expand_rows_full(ctx, instrs, pcs, regs, n_rows, final_regs, final_pc, log_n, d_cols);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return expand_error(ctx);
}

// Many independent small proofs (BASELINE config 4).  A tiny proof is launch-latency bound (about 150 dependent launches),
// so the batch is spread over up to 8 worker contexts of the same device, each with its own stream and host thread:
// the GPU overlaps their kernels.  Proof i is bit-identical to zkir_b200_prove(traces[i]) on a single context.
int zkir_b200_prove_batch(zkir_ctx* ctx, const zkir_params* p, const uint32_t* const* traces, const uint32_t* log_ns,
                          const uint32_t* const* pvs, const uint32_t* const* ios, const size_t* n_ios, uint32_t n_proofs, uint8_t** proofs,
                          size_t* proof_lens) {
  if (!ctx || !traces || !log_ns || !pvs || !ios || !n_ios || !proofs || !proof_lens) return ZKIR_ERR_ARG;
  ctx->err.clear();
  for (uint32_t i = 0; i < n_proofs; i++) { proofs[i] = nullptr; proof_lens[i] = 0; }
  uint32_t nw = n_proofs < 8 ? n_proofs : 8;
  const char* env = getenv("ZKIR_BATCH_WORKERS");
  if (env && atoi(env) > 0) nw = (uint32_t)atoi(env) < n_proofs ? (uint32_t)atoi(env) : n_proofs;
  if (nw <= 1) {
    for (uint32_t i = 0; i < n_proofs; i++) {
      int rc = zkir_b200_set_io(ctx, ios[i], n_ios[i]);
      if (!rc) rc = zkir_b200_prove(ctx, p, traces[i], log_ns[i], pvs[i], &proofs[i], &proof_lens[i]);
      if (rc) { for (uint32_t j = 0; j < i; j++) { free(proofs[j]); proofs[j] = nullptr; } return rc; }
    }
    return 0;
  }
  while (ctx->workers.size() < nw) {
    zkir_ctx* w = nullptr;
    int rc = zkir_b200_create(&w, ctx->device);
    if (rc) { ctx->err = "worker context: " + g_last_error; return rc; }
    if ((rc = zkir_b200_set_program(w, ctx->code.data(), ctx->code.size())) != 0) { ctx->err = "worker context: program"; zkir_b200_destroy(w); return rc; }
    ctx->workers.push_back(w);
  }
  std::vector<int> rcs(nw, 0);
  std::vector<std::thread> th;
  for (uint32_t w = 0; w < nw; w++) {
    th.emplace_back([&, w]() {
      zkir_ctx* wc = ctx->workers[w];
      for (uint32_t i = w; i < n_proofs; i += nw) {
        int rc = zkir_b200_set_io(wc, ios[i], n_ios[i]);
        if (!rc) rc = zkir_b200_prove(wc, p, traces[i], log_ns[i], pvs[i], &proofs[i], &proof_lens[i]);
        if (rc) { rcs[w] = rc; return; }
      }
    });
  }
  for (auto& t : th) t.join();
  for (uint32_t w = 0; w < nw; w++) {
    if (rcs[w]) {
      ctx->err = ctx->workers[w]->err;
      for (uint32_t j = 0; j < n_proofs; j++) { free(proofs[j]); proofs[j] = nullptr; proof_lens[j] = 0; }
      return rcs[w];
    }
  }
  return 0;
}

// ---------------------------------------------------------------- one proof sharded over several GPUs (BASELINE config 5)
int zkir_b200_comm_unique_id(uint8_t id[ZKIR_COMM_ID_LEN]) {
  if (!id) return ZKIR_ERR_ARG;
  return comm_unique_id(id, &g_last_error) == 0 ? 0 : ZKIR_ERR_NCCL;
}

int zkir_b200_comm_init(zkir_ctx* ctx, const uint8_t id[ZKIR_COMM_ID_LEN], int rank, int world) {
  if (!ctx || !id) return ZKIR_ERR_ARG;
  ctx->err.clear();
  if (world < 1 || world > 64 || (world & (world - 1)) || rank < 0 || rank >= world) { ctx->err = "comm_init: world must be a power of two <= 64, 0 <= rank < world"; return ZKIR_ERR_ARG; }
  cudaSetDevice(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  peers_close(ctx);
  comm_destroy(ctx->comm);
  ctx->comm = nullptr; ctx->shards = 1; ctx->shard_lo = 0; ctx->shard_hi = 1;
  if (world == 1) return 0;
  if (comm_create(&ctx->comm, id, rank, world, &ctx->err) != 0) return ZKIR_ERR_NCCL;
  ctx->shards = (u32)world; ctx->shard_lo = (u32)rank; ctx->shard_hi = (u32)rank + 1;
  const char* env = getenv("ZKIR_SHARD_MIN_SEG");
  if (env && atoll(env) > 0) ctx->shard_min_seg = (u64)atoll(env);
  return 0;
}

int zkir_b200_comm_shutdown(zkir_ctx* ctx) {
  if (!ctx) return ZKIR_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->comm && ctx->ws.valid) {   // collective: every rank closes its peer mappings before any rank may free its workspace
    peers_close(ctx);
    int brc = peers_barrier(ctx); if (brc) return brc;
    CU(cudaStreamSynchronize(ctx->stream));
  }
  peers_close(ctx);
  comm_destroy(ctx->comm);
  ctx->comm = nullptr; ctx->shards = 1; ctx->shard_lo = 0; ctx->shard_hi = 1;
  return 0;
}

int zkir_b200_shard_plan(uint32_t world, uint32_t rank, const zkir_params* p, uint32_t log_n, uint64_t min_segment_leaves, uint64_t out[8]) {
  if (!p || !out || world < 1 || world > ZKIR_MAX_SHARDS || (world & (world - 1)) || rank >= world || log_n < 2 || log_n + p->log_blowup > 27) return ZKIR_ERR_ARG;
  FastPlan a, b;
  const ShardPlan sp = shard_plan_for(world, p->width, log_n, p->log_blowup, fast_path_ok(log_n, p->log_blowup, &a, &b),
                                      min_segment_leaves ? min_segment_leaves : 4096);
  const u64 N = 1ull << log_n, M = N << p->log_blowup;
  out[0] = sp.on ? 1 : 0;
  out[1] = sp.on ? sp.c_lo(rank) : 0;          out[2] = sp.on ? sp.c_hi(rank) : p->width;   // trace columns transformed by this rank
  out[3] = sp.on ? (u64)rank * sp.nj : 0;      out[4] = sp.on ? sp.nj : N;                  // points j of every coset = its row segment
  out[5] = sp.on ? sp.p_lo(rank) : 0;          out[6] = sp.on ? sp.p_hi(rank) : 4;          // quotient planes transformed by this rank
  out[7] = sp.on ? M / world : M;                                                           // Merkle leaves per segment
  return 0;
}

int zkir_b200_emulate_shards(zkir_ctx* ctx, uint32_t shards, uint64_t min_segment_leaves) {
  if (!ctx || ctx->comm || shards < 1 || shards > 64 || (shards & (shards - 1))) return ZKIR_ERR_ARG;
  ctx->shards = shards; ctx->shard_lo = 0; ctx->shard_hi = shards;
  if (min_segment_leaves) ctx->shard_min_seg = min_segment_leaves;
  return 0;
}

int zkir_b200_last_stage_ms(zkir_ctx* ctx, float* out) {
  if (!ctx || !ctx->have_stage) return ZKIR_ERR_ARG;
  memcpy(out, ctx->stage_ms, sizeof(ctx->stage_ms));
  return 0;
}
uint64_t zkir_b200_kernel_launches(const zkir_ctx* ctx) { return ctx ? ctx->launches : 0; }
int zkir_b200_timer_start(zkir_ctx* ctx) {
  if (!ctx) return ZKIR_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(cudaEventRecord(ctx->tev[0], ctx->stream));
  return 0;
}
int zkir_b200_timer_stop(zkir_ctx* ctx, float* ms) {
  if (!ctx || !ms) return ZKIR_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(cudaEventRecord(ctx->tev[1], ctx->stream));
  CU(cudaEventSynchronize(ctx->tev[1]));
  CU(cudaEventElapsedTime(ms, ctx->tev[0], ctx->tev[1]));
  return 0;
}

// ---------------------------------------------------------------- per-kernel entry points (canonical device data)
int zkir_b200_ntt(zkir_ctx* ctx, uint32_t* d_cols, uint32_t n_cols, uint32_t log_n, int inverse, uint32_t coset_shift) {
  if (!ctx || !d_cols || log_n < 1 || log_n > 27 || (inverse && coset_shift)) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  const u64 n = 1ull << log_n;
  FastPlan fp;
  if (!getenv("ZKIR_FORCE_GENERIC_NTT") && fast_plan((int)log_n, &fp) && fp.nd == 2) {
    // register-tile kernels: pass A (strided, top digit) into a scratch matrix, pass B (rows, transposing store) back
    u32* tmp = nullptr;
    RC(ensure_scratch(ctx, 0, (size_t)n_cols * n * 4, &tmp));
    RC(fast_ntt_natural(ctx->fast, (int)log_n, inverse != 0, coset_shift % BB_P, d_cols, n, tmp, d_cols, n, n_cols, ctx->stream));
    return 0;
  }
  RC(ensure_ntt_tmp(ctx, n));
  const u32* in_scale = nullptr;
  if (coset_shift) { in_scale = ntt_powers_table(ctx->tables, coset_shift % BB_P, 1, n); if (!in_scale) return ZKIR_ERR_OOM; }
  const u32 ninv = bb_to_mont_c(hinv((u32)(n % BB_P)));
  RC(ntt_run(ctx->tables, d_cols, n, d_cols, n, ctx->ntt_tmp, ctx->ntt_tmp_words, n_cols, log_n, inverse != 0, 0, in_scale, nullptr,
             ninv, inverse != 0, ctx->stream));
  return 0;
}

int zkir_b200_lde(zkir_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, uint32_t n_cols, uint32_t log_n, uint32_t log_blowup) {
  if (!ctx || !d_in || !d_out || log_n < 1 || log_blowup < 1 || log_n + log_blowup > 27) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  const u64 N = 1ull << log_n, M = N << log_blowup;
  FastPlan fp;
  if (!getenv("ZKIR_FORCE_GENERIC_NTT") && fast_plan((int)log_n, &fp)) {
    // same kernels as the prover: digit-reversed coefficients, coset-major LDE, then reorder to the natural order this entry point documents
    u32 *coef = nullptr, *cm = nullptr;
    RC(ensure_scratch(ctx, 0, (size_t)n_cols * N * 4, &coef));
    RC(ensure_scratch(ctx, 1, (size_t)n_cols * M * 4, &cm));
    int frc = fast_intt(ctx->fast, fp, d_in, N, coef, N, n_cols, hinv((u32)(N % BB_P)), nullptr, 32, 0, 0, nullptr, 0, ctx->stream);
    if (!frc) frc = fast_coset_ntt(ctx->fast, fp, coef, N, cm, M, n_cols, 1u << log_blowup, ZKIR_BB_GEN, ZKIR_BB_ROOTS[log_n + log_blowup], 1u, ctx->stream);
    if (!frc) frc = launch_coset_reorder(cm, d_out, n_cols, log_n, log_blowup, 1, ctx->stream, &ctx->launches);
    RC(frc);
    return 0;
  }
  RC(ensure_ntt_tmp(ctx, M));
  const u32* sc = ntt_powers_table(ctx->tables, ZKIR_BB_GEN, hinv((u32)(N % BB_P)), N);
  if (!sc) return ZKIR_ERR_OOM;
  u32* coef = nullptr;
  CU(cudaMallocAsync(&coef, (size_t)n_cols * N * 4, ctx->stream));
  int rc = ntt_run(ctx->tables, d_in, N, coef, N, ctx->ntt_tmp, ctx->ntt_tmp_words, n_cols, log_n, true, 0, nullptr, sc, BB_ONE, false, ctx->stream);
  if (!rc) rc = ntt_run(ctx->tables, coef, N, d_out, M, ctx->ntt_tmp, ctx->ntt_tmp_words, n_cols, log_n + log_blowup, false, log_blowup, nullptr, nullptr,
                        BB_ONE, false, ctx->stream);
  cudaFreeAsync(coef, ctx->stream);
  RC(rc);
  return 0;
}

int zkir_b200_poseidon2_permute(zkir_ctx* ctx, uint32_t* d_states, uint64_t n) {
  if (!ctx || !d_states) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  RC(launch_permute(d_states, n, true, ctx->stream, &ctx->launches));
  return 0;
}

int zkir_b200_merkle_commit(zkir_ctx* ctx, const uint32_t* d_matrix, uint32_t n_cols, uint32_t log_rows, uint32_t* d_tree, uint32_t root[8]) {
  if (!ctx || !d_matrix || !d_tree || !root || n_cols == 0 || log_rows > 27) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  const u64 rows = 1ull << log_rows;
  u32* mont = nullptr;
  CU(cudaMallocAsync(&mont, (size_t)n_cols * rows * 4, ctx->stream));
  int rc = launch_map(mont, d_matrix, (u64)n_cols * rows, 1, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_leaf_hash(mont, rows, n_cols, rows, 0, d_tree, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_merkle_levels(d_tree, rows, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_map(d_tree, d_tree, (2 * rows - 1) * 8, 0, ctx->stream, &ctx->launches);
  cudaFreeAsync(mont, ctx->stream);
  RC(rc);
  CU(cudaMemcpyAsync(root, d_tree + (2 * rows - 2) * 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int zkir_b200_quotient(zkir_ctx* ctx, const zkir_params* p, const uint32_t* d_lde, const uint32_t* d_publde, uint32_t log_n, const uint32_t* pv,
                       const uint32_t lookup[8], const uint32_t alpha[4], uint32_t* d_q) {
  if (!ctx || !p || !d_lde || !d_publde || !pv || !lookup || !alpha || !d_q) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  if (!profile_known(p->width) || p->num_public != ZKIR_NUM_PUBLIC_VALUES || log_n < 2 || p->log_blowup < 1 || log_n + p->log_blowup > 27) { ctx->err = "bad params"; return ZKIR_ERR_ARG; }
  const u32 log_m = log_n + p->log_blowup, AW = profile_aux_width(p->width), PW = profile_pub_width(p->width), WA = p->width + AW, NP = ZKIR_NUM_PUBLIC_VALUES;
  const u64 M = 1ull << log_m;
  u32 *lde_m = nullptr, *lde_cm = nullptr, *pub_m = nullptr, *pub_cm = nullptr, *xs = nullptr, *dinv = nullptr, *small = nullptr;
  CU(cudaMallocAsync(&lde_m, (size_t)WA * M * 4, ctx->stream));
  CU(cudaMallocAsync(&lde_cm, (size_t)WA * M * 4, ctx->stream));
  CU(cudaMallocAsync(&pub_m, (size_t)PW * M * 4, ctx->stream));
  CU(cudaMallocAsync(&pub_cm, (size_t)PW * M * 4, ctx->stream));
  CU(cudaMallocAsync(&xs, M * 4, ctx->stream));
  CU(cudaMallocAsync(&dinv, M * 4, ctx->stream));
  // small: pv[8 (NP padded)] alpha[4] lookup[12: z, theta, sio] apow[K][4] -- the ext4 arrays must stay 16-byte aligned
  CU(cudaMallocAsync(&small, (24 + 4 * ZKIR_PROFILE_MAX_CONSTRAINTS) * 4, ctx->stream));
  u32 h[24] = {0};
  for (u32 i = 0; i < NP; i++) h[i] = bb_to_mont_c(pv[i] % BB_P);
  for (int i = 0; i < 4; i++) h[8 + i] = bb_to_mont_c(alpha[i] % BB_P);
  for (int i = 0; i < 8; i++) h[12 + i] = bb_to_mont_c(lookup[i] % BB_P);
  CU(cudaMemcpyAsync(small, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  RC(launch_io_sum(ctx->d_io, (u32)(ctx->io.size() / 4), small + 12, ctx->stream, &ctx->launches));
  int rc = launch_map(lde_m, d_lde, (u64)WA * M, 1, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_coset_reorder(lde_m, lde_cm, WA, log_n, p->log_blowup, 0, ctx->stream, &ctx->launches);  // the kernel sweeps coset-major rows
  if (!rc) rc = launch_map(pub_m, d_publde, (u64)PW * M, 1, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_coset_reorder(pub_m, pub_cm, PW, log_n, p->log_blowup, 0, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_domain_tables(xs, dinv, log_n, p->log_blowup, ZKIR_BB_GEN, ctx->stream, &ctx->launches);
  QuotientArgs qa;
  qa.lde = lde_cm; qa.publde = pub_cm; qa.q = d_q; qa.log_n = log_n; qa.log_blowup = p->log_blowup; qa.pv = small; qa.alpha = small + 8; qa.lookup = small + 12;
  qa.xs = xs; qa.dinv = dinv; qa.apow_scratch = small + 24;
  if (!rc) rc = profile_is_full(p->width) ? launch_quotient_full(qa, ctx->stream, &ctx->launches) : launch_quotient(qa, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_map(d_q, d_q, 4 * M, 0, ctx->stream, &ctx->launches);
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFreeAsync(lde_m, ctx->stream); cudaFreeAsync(lde_cm, ctx->stream); cudaFreeAsync(pub_m, ctx->stream); cudaFreeAsync(pub_cm, ctx->stream);
  cudaFreeAsync(xs, ctx->stream); cudaFreeAsync(dinv, ctx->stream); cudaFreeAsync(small, ctx->stream);
  RC(rc);
  return 0;
}

// the LogUp aux columns of a canonical device trace for GIVEN lookup challenges (the prover draws them from the transcript)
int zkir_b200_aux_columns(zkir_ctx* ctx, const uint32_t* d_trace, uint32_t log_n, const uint32_t lookup[8], uint32_t* d_aux) {
  if (!ctx || !d_trace || !lookup || !d_aux || log_n < ZKIR_RANGE_BITS || log_n > 26) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  if (ctx->program_version == 0 || ctx->code.size() > (1ull << log_n)) { ctx->err = "no program set, or it does not fit the trace"; return ZKIR_ERR_ARG; }
  const u64 N = 1ull << log_n;
  std::vector<u32> hp((size_t)ZKIR_PROFILE_CORE_PUB * N, 0u);   // per-kernel entry point: core profile
  zkir_public_columns(ZKIR_PROFILE_CORE_WIDTH, log_n, ctx->code.data(), ctx->code.size(), hp.data());
  u32 *pub = nullptr, *lk = nullptr; E4 *rt = nullptr, *bt = nullptr;
  CU(cudaMalloc(&pub, hp.size() * 4)); CU(cudaMalloc(&lk, 48)); CU(cudaMalloc(&rt, N * sizeof(E4))); CU(cudaMalloc(&bt, aux_gen_blocks(N) * sizeof(E4)));
  u32 hl[8];
  for (int i = 0; i < 8; i++) hl[i] = bb_to_mont_c(lookup[i] % BB_P);
  CU(cudaMemcpyAsync(pub, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice, ctx->stream));   // on the stream the kernels run on (see zkir_b200_set_io)
  CU(cudaMemcpyAsync(lk, hl, 32, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_err, 0xff, 16, ctx->stream));
  RC(launch_io_sum(ctx->d_io, (u32)(ctx->io.size() / 4), lk, ctx->stream, &ctx->launches));
  AuxArgs aa;
  aa.trace = d_trace; aa.pub = pub; aa.lookup = lk; aa.aux = d_aux; aa.log_n = log_n; aa.row_tot = rt; aa.blk_tot = bt; aa.err = ctx->d_err + 1;
  int rc = launch_aux_gen(aa, ctx->stream, &ctx->launches);
  CU(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFree(pub); cudaFree(lk); cudaFree(rt); cudaFree(bt);
  RC(rc);
  return expand_error(ctx);
}

int zkir_b200_fri_fold(zkir_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, uint32_t log_n, uint32_t shift, const uint32_t beta[4]) {
  if (!ctx || !d_in || !d_out || log_n < 1 || log_n > 27 || shift == 0) return ZKIR_ERR_ARG;
  ctx->err.clear();
  cudaSetDevice(ctx->device);
  const u64 n = 1ull << log_n, h = n / 2;
  const u32* inv_w = ntt_powers_table(ctx->tables, hinv(ZKIR_BB_ROOTS[log_n]), 1, h);
  if (!inv_w) return ZKIR_ERR_OOM;
  u32 *in_m = nullptr, *b = nullptr;
  CU(cudaMallocAsync(&in_m, n * 16, ctx->stream));
  CU(cudaMallocAsync(&b, 16, ctx->stream));
  u32 hb[4];
  for (int i = 0; i < 4; i++) hb[i] = bb_to_mont_c(beta[i] % BB_P);
  CU(cudaMemcpyAsync(b, hb, 16, cudaMemcpyHostToDevice, ctx->stream));
  int rc = launch_map(in_m, d_in, n * 4, 1, ctx->stream, &ctx->launches);
  if (!rc) rc = launch_fri_fold(reinterpret_cast<const E4*>(in_m), reinterpret_cast<E4*>(d_out), h, b, inv_w, 1,
                                bb_to_mont_c(hinv(hmul(2, shift % BB_P))), ctx->stream, &ctx->launches);
  if (!rc) rc = launch_map(d_out, d_out, h * 4, 0, ctx->stream, &ctx->launches);
  CU(cudaStreamSynchronize(ctx->stream));
  cudaFreeAsync(in_m, ctx->stream); cudaFreeAsync(b, ctx->stream);
  RC(rc);
  return 0;
}

int zkir_b200_dev_alloc(zkir_ctx* ctx, void** d_ptr, size_t bytes) {
  if (!ctx || !d_ptr) return ZKIR_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(cudaMalloc(d_ptr, bytes ? bytes : 16));
  return 0;
}
int zkir_b200_dev_free(zkir_ctx* ctx, void* d_ptr) { if (!ctx) return ZKIR_ERR_ARG; cudaSetDevice(ctx->device); CU(cudaFree(d_ptr)); return 0; }
int zkir_b200_h2d(zkir_ctx* ctx, void* d, const void* h, size_t bytes) {
  if (!ctx) return ZKIR_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int zkir_b200_d2h(zkir_ctx* ctx, void* h, const void* d, size_t bytes) {
  if (!ctx) return ZKIR_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int zkir_b200_sync(zkir_ctx* ctx) { if (!ctx) return ZKIR_ERR_ARG; cudaSetDevice(ctx->device); CU(cudaStreamSynchronize(ctx->stream)); return 0; }

}  // extern "C"
