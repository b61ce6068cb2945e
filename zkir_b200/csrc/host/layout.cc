// zkir_b200_proof_size: host arithmetic only (proof_layout.h), so that the verifier needs no GPU code.
#include "zkir_b200.h"
#include "../proof_layout.h"

extern "C" size_t zkir_b200_proof_size(const zkir_params* p, uint32_t log_n) { return p ? zkir::make_layout(p, log_n).total * 4 : 0; }
