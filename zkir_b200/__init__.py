"""zkir_b200 -- B200-native STARK proving backend for the ZKIR v3.4 VM (trace -> proof path only).

`prove()` / `verify()` / `run()` mirror the runtime surface of seceq/zkir; everything heavy happens in
libzkir_b200.so (hand-written sm_100a CUDA behind the C ABI of include/zkir_b200.h)."""
from .isa import Program, assemble, encode, decode, OPCODES  # noqa: F401
from .runtime import (VM, VMConfig, HaltReason, ExecutionResult, RuntimeError_ as RuntimeError, run, prove, verify,  # noqa: F401
                      ProverConfig, Proof, Context, PinnedBuffer)
from . import air_layout  # noqa: F401
