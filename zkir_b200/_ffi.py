"""ctypes binding of libzkir_b200.so (include/zkir_b200.h).  Fails loudly when the library is missing:
there is no Python/CPU fallback for the proving path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libzkir_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_OOM, ERR_VM, ERR_AIR, ERR_VERIFY = 0, -1, -2, -3, -4, -5, -6, -7
STAGES = ["h2d", "lde", "trace_commit", "aux", "quotient", "quotient_commit", "openings", "fri", "queries_d2h"]


class Params(C.Structure):
    _fields_ = [("log_blowup", C.c_uint32), ("num_queries", C.c_uint32), ("pow_bits", C.c_uint32),
                ("width", C.c_uint32), ("num_public", C.c_uint32)]


class MemOp(C.Structure):
    _fields_ = [("address", C.c_uint64), ("value", C.c_uint64), ("timestamp", C.c_uint64),
                ("is_write", C.c_uint8), ("width", C.c_uint8)]


# every symbol include/zkir_b200.h declares: name -> (restype, argtypes)
u32p, u64p, u8p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_void_p
SYMBOLS = {
    "zkir_b200_create": (C.c_int, [C.POINTER(vp), C.c_int]),
    "zkir_b200_destroy": (None, [vp]),
    "zkir_b200_last_error": (C.c_char_p, [vp]),
    "zkir_b200_alloc_pinned": (vp, [C.c_size_t]),
    "zkir_b200_free_pinned": (None, [vp]),
    "zkir_b200_prove": (C.c_int, [vp, C.POINTER(Params), vp, C.c_uint32, u32p, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "zkir_b200_prove_device": (C.c_int, [vp, C.POINTER(Params), vp, C.c_uint32, u32p, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "zkir_b200_set_program": (C.c_int, [vp, vp, C.c_size_t]),
    "zkir_b200_set_io": (C.c_int, [vp, vp, C.c_size_t]),
    "zkir_b200_prove_rows": (C.c_int, [vp, C.POINTER(Params), vp, vp, vp, C.c_uint64, u64p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, u32p,
                                      C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "zkir_b200_expand_rows": (C.c_int, [vp, vp, vp, vp, C.c_uint64, u64p, C.c_uint64, C.c_uint32, vp]),
    "zkir_b200_expand_rows_full": (C.c_int, [vp, vp, vp, vp, C.c_uint64, u64p, C.c_uint64, C.c_uint32, vp]),
    "zkir_b200_prove_writelog": (C.c_int, [vp, C.POINTER(Params), vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, u32p,
                                          C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "zkir_b200_prove_program": (C.c_int, [vp, C.POINTER(Params), vp, C.c_size_t, vp, C.c_size_t, C.c_uint32, vp, C.c_size_t, C.c_uint64, u32p,
                                         C.POINTER(C.c_uint64), u32p, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "zkir_b200_expand_writelog": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_uint64, C.c_uint32, vp]),
    "zkir_b200_prove_batch": (C.c_int, [vp, C.POINTER(Params), C.POINTER(vp), u32p, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), C.c_uint32,
                                       C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "zkir_b200_comm_unique_id": (C.c_int, [vp]),
    "zkir_b200_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    "zkir_b200_comm_shutdown": (C.c_int, [vp]),
    "zkir_b200_emulate_shards": (C.c_int, [vp, C.c_uint32, C.c_uint64]),
    "zkir_b200_shard_plan": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(Params), C.c_uint32, C.c_uint64, u64p]),
    "zkir_b200_free_proof": (None, [vp]),
    "zkir_b200_proof_size": (C.c_size_t, [C.POINTER(Params), C.c_uint32]),
    "zkir_b200_verify": (C.c_int, [C.POINTER(Params), vp, C.c_size_t, u32p, vp, C.c_size_t, vp, C.c_size_t]),
    "zkir_io_digest": (None, [vp, C.c_size_t, vp]),
    "zkir_host_poseidon2_permute": (None, [vp]),
    "zkir_rom_entry": (None, [C.c_uint32, u32p, u32p]),
    "zkir_program_digest": (None, [vp, C.c_size_t, vp]),
    "zkir_b200_ntt": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32]),
    "zkir_b200_lde": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "zkir_b200_poseidon2_permute": (C.c_int, [vp, vp, C.c_uint64]),
    "zkir_b200_merkle_commit": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, vp, u32p]),
    "zkir_b200_quotient": (C.c_int, [vp, C.POINTER(Params), vp, vp, C.c_uint32, u32p, u32p, u32p, vp]),
    "zkir_b200_aux_columns": (C.c_int, [vp, vp, C.c_uint32, u32p, vp]),
    "zkir_b200_fri_fold": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, u32p]),
    "zkir_b200_dev_alloc": (C.c_int, [vp, C.POINTER(vp), C.c_size_t]),
    "zkir_b200_dev_free": (C.c_int, [vp, vp]),
    "zkir_b200_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "zkir_b200_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "zkir_b200_sync": (C.c_int, [vp]),
    "zkir_b200_last_stage_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "zkir_b200_kernel_launches": (C.c_uint64, [vp]),
    "zkir_b200_timer_start": (C.c_int, [vp]),
    "zkir_b200_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "zkir_encode": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32]),
    "zkir_decode": (C.c_int, [C.c_uint32, u32p]),
    "zkir_vm_run": (C.c_int, [u32p, C.c_size_t, u8p, C.c_size_t, C.c_uint32, u64p, C.c_size_t, C.c_uint64, C.c_int, C.POINTER(vp)]),
    "zkir_vm_free": (None, [vp]),
    "zkir_vm_enable_poseidon2": (None, [C.c_int]),
    "zkir_vm_last_error": (C.c_char_p, []),
    "zkir_vm_cycles": (C.c_uint64, [vp]),
    "zkir_vm_halt_kind": (C.c_int, [vp]),
    "zkir_vm_exit_code": (C.c_uint64, [vp]),
    "zkir_vm_num_outputs": (C.c_size_t, [vp]),
    "zkir_vm_outputs": (u64p, [vp]),
    "zkir_vm_trace_len": (C.c_size_t, [vp]),
    "zkir_vm_trace_pc": (u64p, [vp]),
    "zkir_vm_trace_instr": (u32p, [vp]),
    "zkir_vm_trace_regs": (u64p, [vp]),
    "zkir_vm_trace_aux": (u64p, [vp]),
    "zkir_vm_trace_memop_begin": (u64p, [vp]),
    "zkir_vm_trace_memops": (C.POINTER(MemOp), [vp]),
    "zkir_vm_trace_writelog": (C.c_int, [vp, vp, vp]),
    "zkir_vm_run_writelog": (C.c_int, [u32p, C.c_size_t, u8p, C.c_size_t, C.c_uint32, u64p, C.c_size_t, C.c_uint64, vp, vp, vp, C.c_uint64, C.POINTER(vp)]),
    "zkir_vm_run_writelog_cb": (C.c_int, [u32p, C.c_size_t, u8p, C.c_size_t, C.c_uint32, u64p, C.c_size_t, C.c_uint64, vp, vp, vp, C.c_uint64, vp, vp, C.c_uint64, C.POINTER(vp)]),
    "zkir_vm_logged_rows": (C.c_uint64, [vp]),
    "zkir_vm_run_writelog_mem_cb": (C.c_int, [u32p, C.c_size_t, u8p, C.c_size_t, C.c_uint32, u64p, C.c_size_t, C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64, vp, vp,
                                            C.c_uint64, C.POINTER(vp)]),
    "zkir_b200_prove_writelog_mem": (C.c_int, [vp, C.POINTER(Params), vp, vp, vp, vp, vp, C.c_uint64, vp, vp, vp, C.c_size_t, C.c_uint64, C.c_uint32, C.c_uint64,
                                              C.c_int, C.c_uint32, u32p, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "zkir_vm_memlog_count": (C.c_size_t, [vp]),
    "zkir_vm_memlog_widx": (u64p, [vp]),
    "zkir_vm_memlog_word": (u64p, [vp]),
    "zkir_vm_memlog_ts": (u32p, [vp]),
    "zkir_vm_io_len": (C.c_size_t, [vp]),
    "zkir_vm_io": (u32p, [vp]),
    "zkir_vm_poseidon2_count": (C.c_size_t, [vp]),
    "zkir_vm_poseidon2_witness": (u32p, [vp]),
    "zkir_vm_code_len": (C.c_size_t, [vp]),
    "zkir_vm_code": (u32p, [vp]),
    "zkir_vm_final_pc": (C.c_uint64, [vp]),
    "zkir_vm_final_regs": (u64p, [vp]),
    "zkir_pack_min_log_n": (C.c_uint32, [vp]),
    "zkir_pack_trace": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, u32p]),
    "zkir_pack_trace_full": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp, u32p]),
    "zkir_pack_rows": (C.c_int, [vp, vp, vp, C.c_uint64, u64p, C.c_uint64, vp, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, vp, u32p]),
    "zkir_pack_rows_full": (C.c_int, [vp, vp, vp, C.c_uint64, u64p, C.c_uint64, vp, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32, vp, u32p]),
    "zkir_program_profile": (C.c_int, [vp, C.c_size_t]),
    "zkir_public_row": (None, [C.c_uint32, C.c_uint64, vp, C.c_size_t, u32p]),
    "zkir_public_rows": (C.c_uint64, [C.c_uint32, C.c_size_t]),
    "zkir_image_words": (C.c_uint64, [C.c_size_t]),
    "zkir_public_columns": (None, [C.c_uint32, C.c_uint32, vp, C.c_size_t, vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m zkir_b200.build` (nvcc, sm_100a). "
                "zkir_b200 has no CPU fallback for the proving path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)  # AttributeError if the header and the library ever disagree
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib
