// NCCL bound at run time (dlopen), so that libzkir_b200.so has no link-time dependency on a particular libnccl: inside a
// Python process the copy torch already loaded is reused, a Rust/C++ host gets the system one.  Only the collectives the
// sharded prover needs (docs/PROVER_SPEC.md section 6: all-gather of Merkle segment roots and row-sharded planes,
// all-reduce of the disjoint query pieces).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <string>

namespace zkir {

struct Comm;  // one NCCL communicator (one rank = one GPU)

#define ZKIR_COMM_ID_BYTES 128
int comm_unique_id(unsigned char id[ZKIR_COMM_ID_BYTES], std::string* err);
int comm_create(Comm** out, const unsigned char id[ZKIR_COMM_ID_BYTES], int rank, int world, std::string* err);
void comm_destroy(Comm*);
// in place: every rank contributes words [rank * words_per_rank, +words_per_rank) of buf
int comm_all_gather_u32(Comm*, unsigned* buf, size_t words_per_rank, cudaStream_t st, std::string* err);
// in place, sum of uint32 (used on buffers whose non-zero pieces are disjoint across ranks: sum == union)
int comm_all_reduce_sum_u32(Comm*, unsigned* buf, size_t words, cudaStream_t st, std::string* err);
// several all-gathers fused into one NCCL group (one launch): piece j is bufs[j] with words_per_rank[j]
int comm_all_gather_group_u32(Comm*, unsigned* const* bufs, const size_t* words_per_rank, int n, cudaStream_t st, std::string* err);

// point-to-point exchange, all operations in ONE NCCL group; for every ordered pair of ranks the sends of the source must be listed
// in the same order as the matching receives of the destination
struct P2POp { int peer; int is_send; unsigned* ptr; size_t words; };
int comm_exchange_u32(Comm*, const P2POp* ops, size_t n, cudaStream_t st, std::string* err);
int comm_rank(const Comm*);
int comm_world(const Comm*);

}  // namespace zkir
