// Selects the AIR profile of a translation unit (docs/PROVER_SPEC.md section 3.7).  The files that instantiate the generated
// constraint list -- quotient.cu, aux_gen.cu, host/pack.cc, host/verify.cc, and oracle/oracle.cc -- are compiled once per profile;
// the `full` build defines ZKIR_PROFILE_FULL and its entry points carry the suffix _full (ZKIR_PF).  The generated headers of the two
// profiles define the same names (ZKIR_AIR_*, ZKIR_COL_*, zkir_air_eval), so a translation unit includes exactly one of them; code
// that serves both profiles (prover.cu, proof_layout.h) uses air_profiles_generated.h and picks by zkir_params.width.
#pragma once
#ifdef ZKIR_PROFILE_FULL
#include "air_generated_full.h"
#include "air_columns_full.h"
#define ZKIR_PF(name) name##_full
#else
#include "air_generated.h"
#include "air_columns.h"
#define ZKIR_PF(name) name
#endif
