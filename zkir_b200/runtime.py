"""Host-side mirror of the reference's runtime interface for the Program -> Proof path.

Same names and argument meaning as zkir-runtime's public API (zkir-runtime/src/lib.rs:29-62): `VMConfig`
(vm.rs:15-50), `VM(program, inputs, config).run()` / `run(program, inputs)` returning an `ExecutionResult`
(vm.rs:54-78), `HaltReason` (state.rs), plus the `prove()` the north star places next to `run()`.
The interpreter is the C++ restatement behind the C ABI (zkir_b200/csrc/host/vm.cc); proving always goes to the
CUDA library -- there is no CPU fallback here.
"""
import ctypes as C
from dataclasses import dataclass, field
import numpy as np
from . import _ffi
from .air_layout import WIDTH, NUM_PUBLIC, AUX_WIDTH, PUB_WIDTH, MIN_LOG_N
from . import air_layout_full as _full_layout

FULL_WIDTH = _full_layout.WIDTH


def program_profile(program):
    """AIR profile of a program (docs/PROVER_SPEC.md section 3.7): "core" if it stays inside the 18 core opcodes, "full" if it uses MUL /
    DIV / bitwise / shift / signed-compare opcodes, None if it contains an opcode no profile constrains (the loads and stores)."""
    code = np.ascontiguousarray(getattr(program, "code", program), dtype=np.uint32)
    return {1: "core", 0: "full", -1: None}[_ffi.lib().zkir_program_profile(code.ctypes.data, int(code.shape[0]))]


def profile_width(profile):
    return FULL_WIDTH if profile == "full" else WIDTH


P = 2013265921


class RuntimeError_(Exception):
    """zkir-runtime/src/error.rs:6-39 (`RuntimeError`); `.code` is the C ABI return value."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


@dataclass
class VMConfig:  # vm.rs:15-50
    max_cycles: int = 1_000_000
    trace: bool = False
    enable_range_checking: bool = False
    enable_execution_trace: bool = False
    enable_deferred_model: bool = False
    enable_poseidon2_syscall: bool = False   # not upstream: syscall 4 is a stub that errors there (crypto.rs:299-315)


@dataclass
class HaltReason:  # Exit(code) | Ebreak | CycleLimit
    kind: str
    code: int = 0

    def __eq__(self, other):
        return isinstance(other, HaltReason) and self.kind == other.kind and (self.kind != "Exit" or self.code == other.code)

    @staticmethod
    def Exit(code):
        return HaltReason("Exit", code)


HaltReason.Ebreak = HaltReason("Ebreak")
HaltReason.CycleLimit = HaltReason("CycleLimit")
_HALT = {0: "Exit", 1: "Ebreak", 2: "CycleLimit"}


@dataclass
class MemoryOp:  # zkir-spec/src/trace.rs:149-167
    address: int
    value: int
    timestamp: int
    is_write: bool
    width: int


@dataclass
class TraceRow:  # zkir-spec/src/trace.rs:24-50 (bounds / register_states are host bookkeeping, not recorded)
    cycle: int
    pc: int
    instruction: int
    registers: list
    memory_ops: list


_HALT_KIND = {"Exit": 0, "Ebreak": 1, "CycleLimit": 2}


class ExecutionResult:  # vm.rs:54-78
    def __init__(self, handle, entry_point, program=None, wl_arrays=None):
        self._h = handle
        self._entry = entry_point
        self.program = program
        self._wl_arrays = wl_arrays   # write-log mode: the (pinned) arrays the interpreter recorded into
        l = _ffi.lib()
        self.cycles = l.zkir_vm_cycles(handle)
        n = l.zkir_vm_num_outputs(handle)
        self.outputs = [l.zkir_vm_outputs(handle)[i] for i in range(n)]
        self.halt_reason = HaltReason(_HALT[l.zkir_vm_halt_kind(handle)], l.zkir_vm_exit_code(handle))
        self.range_check_witnesses = []
        self.normalization_witnesses = []
        self._trace = None

    def __del__(self):
        if getattr(self, "_h", None):
            _ffi.lib().zkir_vm_free(self._h)
            self._h = None

    @property
    def io(self):
        """public I/O transcript of the run: uint32 [n_events, 4] = (cycle, kind 0 READ / 1 WRITE, value lo20, value hi20)"""
        l = _ffi.lib()
        n = l.zkir_vm_io_len(self._h)
        return np.ctypeslib.as_array(l.zkir_vm_io(self._h), shape=(n, 4)).copy() if n else np.zeros((0, 4), dtype=np.uint32)

    @property
    def poseidon2_witness(self):
        """Poseidon2Witness records of a traced run (zkir-spec/src/trace.rs:287-304): (timestamps uint64 [n], input_state uint32 [n, 16],
        output_state uint32 [n, 16]), one per SYS_POSEIDON2 call."""
        l = _ffi.lib()
        n = l.zkir_vm_poseidon2_count(self._h)
        if n == 0:
            return np.zeros(0, dtype=np.uint64), np.zeros((0, 16), dtype=np.uint32), np.zeros((0, 16), dtype=np.uint32)
        w = np.ctypeslib.as_array(l.zkir_vm_poseidon2_witness(self._h), shape=(n, 34)).copy()
        return w[:, 0].astype(np.uint64) | (w[:, 1].astype(np.uint64) << np.uint64(32)), w[:, 2:18].copy(), w[:, 18:34].copy()

    @property
    def trace_len(self):
        return _ffi.lib().zkir_vm_trace_len(self._h)

    @property
    def execution_trace(self):
        if self._trace is None:
            l = _ffi.lib()
            n = l.zkir_vm_trace_len(self._h)
            pcs, ins, regs = l.zkir_vm_trace_pc(self._h), l.zkir_vm_trace_instr(self._h), l.zkir_vm_trace_regs(self._h)
            beg, ops = l.zkir_vm_trace_memop_begin(self._h), l.zkir_vm_trace_memops(self._h)
            rows = []
            for i in range(n):
                mops = [MemoryOp(ops[j].address, ops[j].value, ops[j].timestamp, bool(ops[j].is_write), ops[j].width)
                        for j in range(beg[i], beg[i + 1])]
                rows.append(TraceRow(i, pcs[i], ins[i], [regs[16 * i + k] for k in range(16)], mops))
            self._trace = rows
        return self._trace

    def get_memory_trace(self):  # vm.rs:85-94, order trace.rs:210-223
        ops = [op for row in self.execution_trace for op in row.memory_ops]
        return sorted(ops, key=lambda o: (o.timestamp, o.address, o.is_write))

    def memory_op_count(self):
        return sum(len(r.memory_ops) for r in self.execution_trace)

    # ---- raw rows, as `Vec<TraceRow>` holds them (trace.rs:24-50): zero-copy numpy views of the recorder's arrays
    def rows(self):
        """-> dict(pcs u64[T], instrs u32[T], regs u64[T,16], final_regs u64[16], final_pc, exit_code)."""
        l = _ffi.lib()
        n = l.zkir_vm_trace_len(self._h)
        as_np = lambda ptr, shape, dt: np.ctypeslib.as_array(ptr, shape=shape).view(dt) if n else np.zeros(shape, dtype=dt)
        return {
            "pcs": as_np(l.zkir_vm_trace_pc(self._h), (n,), np.uint64),
            "instrs": as_np(l.zkir_vm_trace_instr(self._h), (n,), np.uint32),
            "regs": as_np(l.zkir_vm_trace_regs(self._h), (n * 16,), np.uint64).reshape(n, 16),
            "final_regs": np.ctypeslib.as_array(l.zkir_vm_final_regs(self._h), shape=(16,)).copy(),
            "final_pc": l.zkir_vm_final_pc(self._h),
            "exit_code": self.halt_reason.code if self.halt_reason.kind == "Exit" else 0,
            "halt_kind": _HALT_KIND[self.halt_reason.kind],
            "entry_point": self._entry,
            "program": self.program, "io": self.io,
        }

    def writelog(self, out=None):
        """Register write log of the recorded rows (zkir_vm_trace_writelog): dict(pcs u32[T], instrs u32[T], wlog u64[T], ...).
        `out` = optional dict of preallocated (pinned) arrays 'pcs', 'wlog'."""
        l = _ffi.lib()
        if self._wl_arrays is not None:   # recorded directly by the interpreter (zkir_vm_run_writelog): nothing to convert
            n = int(l.zkir_vm_logged_rows(self._h))
            a = self._wl_arrays
            wl = {"pcs": a["pcs"][:n], "instrs": a["instrs"][:n], "wlog": a["wlog"][:n], "final_pc": l.zkir_vm_final_pc(self._h),
                  "exit_code": self.halt_reason.code if self.halt_reason.kind == "Exit" else 0,
                  "halt_kind": _HALT_KIND[self.halt_reason.kind], "entry_point": self._entry, "program": self.program, "io": self.io}
            if "mem_old" in a:   # memory log (full profile): per-row arrays + the touched words, ascending
                nw = int(l.zkir_vm_memlog_count(self._h))
                view = lambda ptr, dt: np.ctypeslib.as_array(ptr, shape=(nw,)).copy() if nw else np.zeros(0, dtype=dt)
                wl.update(mem_old=a["mem_old"][:n], mem_pts=a["mem_pts"][:n], mem_widx=view(l.zkir_vm_memlog_widx(self._h), np.uint64),
                          mem_word=view(l.zkir_vm_memlog_word(self._h), np.uint64), mem_ts=view(l.zkir_vm_memlog_ts(self._h), np.uint32))
            return wl
        n = l.zkir_vm_trace_len(self._h)
        pcs = out["pcs"] if out else np.empty(n, dtype=np.uint32)
        wlog = out["wlog"] if out else np.empty(n, dtype=np.uint64)
        assert pcs.dtype == np.uint32 and wlog.dtype == np.uint64 and pcs.shape == (n,) and wlog.shape == (n,)
        rc = l.zkir_vm_trace_writelog(self._h, pcs.ctypes.data, wlog.ctypes.data)
        if rc != 0:
            raise RuntimeError_(rc, l.zkir_vm_last_error().decode())
        r = self.rows()
        return {"pcs": pcs, "instrs": r["instrs"], "wlog": wlog, "final_pc": r["final_pc"], "exit_code": r["exit_code"],
                "halt_kind": r["halt_kind"], "entry_point": r["entry_point"], "program": self.program, "io": self.io}

    # ---- trace -> columns ("converter", trace.rs:41)
    def min_log_n(self):
        return _ffi.lib().zkir_pack_min_log_n(self._h)

    def pack(self, log_n=None, out=None, profile=None):
        """-> (cols[width][1<<log_n] uint32, public_values[5] uint32).  `out` may be a pinned buffer view.  `profile`: "core" (88 columns),
        "full" (the full-ISA table) or None = what the program needs (program_profile)."""
        l = _ffi.lib()
        if log_n is None:
            log_n = self.min_log_n()
        if profile is None:
            profile = (program_profile(self.program) if self.program is not None else "core") or "core"   # an unconstrained opcode: the core packer names the row
        width = profile_width(profile)
        cols = out if out is not None else np.empty((width, 1 << log_n), dtype=np.uint32)
        assert cols.dtype == np.uint32 and cols.shape == (width, 1 << log_n) and cols.flags["C_CONTIGUOUS"]
        pv = np.zeros(NUM_PUBLIC, dtype=np.uint32)
        fn = l.zkir_pack_trace_full if profile == "full" else l.zkir_pack_trace
        rc = fn(self._h, self._entry, log_n, cols.ctypes.data, pv.ctypes.data_as(_ffi.u32p))
        if rc != 0:
            raise RuntimeError_(rc, l.zkir_b200_last_error(None).decode())
        return cols, pv


class VM:  # vm.rs:104-205
    def __init__(self, program, inputs=(), config=None):
        if program.entry_point < 0x1000:  # vm.rs:141-147 panics
            raise RuntimeError_(_ffi.ERR_ARG, f"Program appears to be in debug format (entry_point={program.entry_point:#x})")
        self.program, self.inputs, self.config = program, list(inputs), config or VMConfig()

    def run(self):
        l = _ffi.lib()
        code = (C.c_uint32 * max(1, len(self.program.code)))(*self.program.code)
        data = (C.c_uint8 * max(1, len(self.program.data)))(*self.program.data)
        inp = (C.c_uint64 * max(1, len(self.inputs)))(*[v & (2**64 - 1) for v in self.inputs])
        h = C.c_void_p()
        l.zkir_vm_enable_poseidon2(int(self.config.enable_poseidon2_syscall))
        rc = l.zkir_vm_run(code, len(self.program.code), data, len(self.program.data), self.program.entry_point,
                           inp, len(self.inputs), self.config.max_cycles, int(self.config.enable_execution_trace), C.byref(h))
        if rc != 0:
            raise RuntimeError_(rc, l.zkir_vm_last_error().decode())
        return ExecutionResult(h, self.program.entry_point, self.program)

    def run_writelog(self, out=None, memory_log=None):
        """Run with the register write log recorded STRAIGHT into `out` = dict(pcs u32[cap], instrs u32[cap], wlog u64[cap])
        (pinned arrays, e.g. PinnedBuffer.array) while the program executes: zkir_vm_run_writelog.  No per-cycle register
        snapshot exists afterwards: `.writelog()` of the result returns views of the arrays.
        `memory_log` (None = when the program needs the full AIR profile): also record mem_old u64[cap] / mem_pts u32[cap], the word before
        each load / store and its previous timestamp (zkir_vm_run_writelog_mem_cb): what Context.prove_writelog needs for such programs."""
        l = _ffi.lib()
        if memory_log is None:
            memory_log = program_profile(self.program) == "full"
        cap = int(self.config.max_cycles) if out is None else int(out["pcs"].shape[0])
        if out is None:
            out = {"pcs": np.empty(cap, dtype=np.uint32), "instrs": np.empty(cap, dtype=np.uint32), "wlog": np.empty(cap, dtype=np.uint64)}
        if memory_log and "mem_old" not in out:
            out = dict(out, mem_old=np.zeros(cap, dtype=np.uint64), mem_pts=np.zeros(cap, dtype=np.uint32))
        assert out["pcs"].dtype == np.uint32 and out["instrs"].dtype == np.uint32 and out["wlog"].dtype == np.uint64
        code = (C.c_uint32 * max(1, len(self.program.code)))(*self.program.code)
        data = (C.c_uint8 * max(1, len(self.program.data)))(*self.program.data)
        inp = (C.c_uint64 * max(1, len(self.inputs)))(*[v & (2**64 - 1) for v in self.inputs])
        h = C.c_void_p()
        l.zkir_vm_enable_poseidon2(int(self.config.enable_poseidon2_syscall))
        if memory_log:
            assert out["mem_old"].dtype == np.uint64 and out["mem_pts"].dtype == np.uint32 and out["mem_old"].shape[0] >= cap and out["mem_pts"].shape[0] >= cap
            rc = l.zkir_vm_run_writelog_mem_cb(code, len(self.program.code), data, len(self.program.data), self.program.entry_point,
                                               inp, len(self.inputs), self.config.max_cycles, out["pcs"].ctypes.data, out["instrs"].ctypes.data,
                                               out["wlog"].ctypes.data, out["mem_old"].ctypes.data, out["mem_pts"].ctypes.data, cap, None, None, 0, C.byref(h))
        else:
            rc = l.zkir_vm_run_writelog(code, len(self.program.code), data, len(self.program.data), self.program.entry_point,
                                        inp, len(self.inputs), self.config.max_cycles, out["pcs"].ctypes.data, out["instrs"].ctypes.data,
                                        out["wlog"].ctypes.data, cap, C.byref(h))
        if rc != 0:
            raise RuntimeError_(rc, l.zkir_vm_last_error().decode())
        return ExecutionResult(h, self.program.entry_point, self.program, wl_arrays=out)


def run(program, inputs=()):  # zkir-runtime/src/lib.rs:59-62
    return VM(program, inputs, VMConfig()).run()


# ------------------------------------------------------------------------------------------------ proving
@dataclass
class ProverConfig:
    log_blowup: int = 1
    num_queries: int = 100
    pow_bits: int = 16
    max_cycles: int = (1 << 24) + 16
    device: int = 0
    enable_poseidon2_syscall: bool = False

    def params(self, width=WIDTH):
        return _ffi.Params(self.log_blowup, self.num_queries, self.pow_bits, width, NUM_PUBLIC)


@dataclass
class Proof:
    bytes_: bytes
    public_values: np.ndarray
    log_n: int
    cycles: int
    outputs: list = field(default_factory=list)
    stage_ms: dict = field(default_factory=dict)
    program: object = None
    io: object = None


class Context:
    """One zkir_ctx (one GPU).  Raises if the CUDA library or a B200 is missing."""

    def __init__(self, device=0):
        self._l = _ffi.lib()
        self._h = C.c_void_p()
        rc = self._l.zkir_b200_create(C.byref(self._h), device)
        if rc != 0:
            raise RuntimeError_(rc, "zkir_b200_create: " + self._l.zkir_b200_last_error(None).decode())

    def close(self):
        if self._h:
            self._l.zkir_b200_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError_(rc, self._l.zkir_b200_last_error(self._h).decode())

    def set_program(self, program):
        """The program whose executions this context proves (zkir_b200_set_program): `Program`, a sequence of code words, or an
        `ExecutionResult` (then its public I/O transcript is set too: the two halves of the statement besides the public values)."""
        if isinstance(program, ExecutionResult):
            self.set_io(program.io)
            program = program.program
        code = np.ascontiguousarray(getattr(program, "code", program), dtype=np.uint32)
        self._check(self._l.zkir_b200_set_program(self._h, code.ctypes.data, int(code.shape[0])))

    def set_io(self, io):
        """The public I/O transcript of the execution about to be proven (zkir_b200_set_io): uint32 [n, 4] as `ExecutionResult.io`."""
        ev = np.ascontiguousarray(io if io is not None else np.zeros((0, 4)), dtype=np.uint32).reshape(-1, 4)
        self._check(self._l.zkir_b200_set_io(self._h, ev.ctypes.data, int(ev.shape[0])))

    # -- device memory helpers
    def alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self._l.zkir_b200_dev_alloc(self._h, C.byref(p), nbytes))
        return p

    def free(self, p):
        self._check(self._l.zkir_b200_dev_free(self._h, p))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.alloc(arr.nbytes)
        self._check(self._l.zkir_b200_h2d(self._h, p, arr.ctypes.data, arr.nbytes))
        return p

    def to_host(self, p, shape, dtype=np.uint32):
        out = np.empty(shape, dtype=dtype)
        self._check(self._l.zkir_b200_d2h(self._h, out.ctypes.data, p, out.nbytes))
        return out

    def sync(self):
        self._check(self._l.zkir_b200_sync(self._h))

    @property
    def kernel_launches(self):
        return self._l.zkir_b200_kernel_launches(self._h)

    def timer_start(self):
        self._check(self._l.zkir_b200_timer_start(self._h))

    def timer_stop(self):
        """ms between timer_start() and now, CUDA events on the library's launch stream."""
        ms = C.c_float()
        self._check(self._l.zkir_b200_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def stage_ms(self):
        """per-stage device times of the last proof; {} when it was a CUDA-graph replay (no per-stage events)"""
        out = (C.c_float * len(_ffi.STAGES))()
        if self._l.zkir_b200_last_stage_ms(self._h, out) != 0:
            return {}
        return dict(zip(_ffi.STAGES, list(out)))

    # -- one proof sharded over several GPUs (include/zkir_b200.h: zkir_b200_comm_*)
    def comm_init(self, rank=None, world=None, group=None):
        """Link this context with the contexts of the other ranks (one process per GPU, torch.distributed already
        initialised: NCCL on the GPU box, gloo works too -- it only carries the 128-byte id).  Afterwards every prove_*
        call on this context is collective (same trace on every rank) and returns the single-GPU proof bytes."""
        from . import multi
        rank, world, ident = multi.exchange_comm_id(self._l, rank, world, group)
        if world > 1:
            self._check(self._l.zkir_b200_comm_init(self._h, ident, rank, world))
        return rank, world

    def comm_shutdown(self):
        self._check(self._l.zkir_b200_comm_shutdown(self._h))

    def emulate_shards(self, shards, min_segment_leaves=0):
        """Test hook: run the sharded code path (segment kernels, no NCCL) for `shards` segments on this one GPU."""
        self._check(self._l.zkir_b200_emulate_shards(self._h, shards, min_segment_leaves))

    # -- hot path
    def prove_columns(self, cols, public_values, cfg, device_resident=None, program=None):
        """cols: host uint32 [WIDTH][2^log_n] (or `device_resident`: a device pointer with the same layout).  `program`: sets the
        context's program first (zkir_b200_set_program; a no-op when it is unchanged)."""
        if program is not None:
            self.set_program(program)
        params = cfg.params(WIDTH if cols is None else int(cols.shape[0]))   # the profile is the table's width (device-resident: pass the host array or None = core)
        pv = np.ascontiguousarray(public_values, dtype=np.uint32)
        proof, plen = C.c_void_p(), C.c_size_t()
        if device_resident is not None:
            ptr, log_n = device_resident
            rc = self._l.zkir_b200_prove_device(self._h, C.byref(params), ptr, log_n, pv.ctypes.data_as(_ffi.u32p), C.byref(proof), C.byref(plen))
        else:
            assert cols.dtype == np.uint32 and cols.flags["C_CONTIGUOUS"] and cols.shape[0] in (WIDTH, FULL_WIDTH)
            log_n = int(cols.shape[1]).bit_length() - 1
            assert cols.shape[1] == 1 << log_n
            rc = self._l.zkir_b200_prove(self._h, C.byref(params), cols.ctypes.data, log_n, pv.ctypes.data_as(_ffi.u32p), C.byref(proof), C.byref(plen))
        self._check(rc)
        out = C.string_at(proof, plen.value)
        self._l.zkir_b200_free_proof(proof)
        return out

    def prove_batch(self, cols_list, pv_list, cfg, program=None, io_list=None):
        """Independent proofs of many (small) traces of ONE program: zkir_b200_prove_batch.  `io_list[i]` = the public I/O transcript of
        execution i.  Returns the list of proof bytes."""
        if program is not None:
            self.set_program(program)
        n = len(cols_list)
        ios = [np.ascontiguousarray(e if e is not None else np.zeros((0, 4)), dtype=np.uint32).reshape(-1, 4) for e in (io_list or [None] * n)]
        iop = (C.c_void_p * n)(*[e.ctypes.data for e in ios])
        ion = (C.c_size_t * n)(*[int(e.shape[0]) for e in ios])
        params = cfg.params()
        keep = [np.ascontiguousarray(c, dtype=np.uint32) for c in cols_list]
        pvs = [np.ascontiguousarray(p, dtype=np.uint32) for p in pv_list]
        tr = (C.c_void_p * n)(*[c.ctypes.data for c in keep])
        pp = (C.c_void_p * n)(*[p.ctypes.data for p in pvs])
        lg = (C.c_uint32 * n)(*[int(c.shape[1]).bit_length() - 1 for c in keep])
        out = (C.c_void_p * n)()
        lens = (C.c_size_t * n)()
        self._check(self._l.zkir_b200_prove_batch(self._h, C.byref(params), tr, lg, pp, iop, ion, n, out, lens))
        res = []
        for i in range(n):
            res.append(C.string_at(out[i], lens[i]))
            self._l.zkir_b200_free_proof(out[i])
        return res

    def prove_rows(self, rows, cfg, log_n=None, profile=None):
        """rows: dict as returned by ExecutionResult.rows() (arrays may live in PinnedBuffers).  Core profile: the device runs the
        converter; full profile (`profile="full"`, or None = what the rows' program needs): the library packs the wide table on the host.
        Returns (proof bytes, public values)."""
        if rows.get("program") is not None:
            self.set_program(rows["program"])
            if profile is None:
                profile = program_profile(rows["program"])
        if rows.get("io") is not None:
            self.set_io(rows["io"])
        params = cfg.params(profile_width(profile))
        n = int(rows["pcs"].shape[0])
        if log_n is None:
            log_n = max(MIN_LOG_N, n.bit_length(), (len(getattr(rows.get("program"), "code", ())) - 1).bit_length())   # one padding row, the ROM
        pcs, ins, regs = (np.ascontiguousarray(rows[k]) for k in ("pcs", "instrs", "regs"))
        assert pcs.dtype == np.uint64 and ins.dtype == np.uint32 and regs.dtype == np.uint64 and regs.shape == (n, 16)
        fr = np.ascontiguousarray(rows["final_regs"], dtype=np.uint64)
        pv = np.zeros(NUM_PUBLIC, dtype=np.uint32)
        proof, plen = C.c_void_p(), C.c_size_t()
        rc = self._l.zkir_b200_prove_rows(self._h, C.byref(params), pcs.ctypes.data, ins.ctypes.data, regs.ctypes.data, n,
                                          fr.ctypes.data_as(_ffi.u64p), int(rows["final_pc"]), int(rows["entry_point"]), int(rows["exit_code"]),
                                          int(rows["halt_kind"]), log_n, pv.ctypes.data_as(_ffi.u32p), C.byref(proof), C.byref(plen))
        self._check(rc)
        out = C.string_at(proof, plen.value)
        self._l.zkir_b200_free_proof(proof)
        return out, pv

    def prove_writelog(self, wl, cfg, log_n=None):
        """wl: dict as returned by ExecutionResult.writelog().  Returns (proof bytes, public values)."""
        if wl.get("program") is not None:
            self.set_program(wl["program"])
        if wl.get("io") is not None:
            self.set_io(wl["io"])
        full = "mem_old" in wl     # write log + memory log: the full profile (zkir_b200_prove_writelog_mem)
        params = cfg.params(FULL_WIDTH if full else WIDTH)
        n = int(wl["pcs"].shape[0])
        if log_n is None:
            log_n = max(MIN_LOG_N, n.bit_length(), (len(getattr(wl.get("program"), "code", ())) - 1).bit_length()) if full else max(MIN_LOG_N, (n - 1).bit_length())
        pcs, ins, wlog = (np.ascontiguousarray(wl[k]) for k in ("pcs", "instrs", "wlog"))
        assert pcs.dtype == np.uint32 and ins.dtype == np.uint32 and wlog.dtype == np.uint64
        pv = np.zeros(NUM_PUBLIC, dtype=np.uint32)
        proof, plen = C.c_void_p(), C.c_size_t()
        if full:
            old, pts = np.ascontiguousarray(wl["mem_old"]), np.ascontiguousarray(wl["mem_pts"])
            widx, word, ts = (np.ascontiguousarray(wl[k]) for k in ("mem_widx", "mem_word", "mem_ts"))
            assert old.dtype == np.uint64 and pts.dtype == np.uint32 and widx.dtype == np.uint64 and word.dtype == np.uint64 and ts.dtype == np.uint32
            rc = self._l.zkir_b200_prove_writelog_mem(self._h, C.byref(params), pcs.ctypes.data, ins.ctypes.data, wlog.ctypes.data, old.ctypes.data, pts.ctypes.data, n,
                                                      widx.ctypes.data, word.ctypes.data, ts.ctypes.data, int(widx.shape[0]), int(wl["final_pc"]),
                                                      int(wl["entry_point"]), int(wl["exit_code"]), int(wl["halt_kind"]), log_n, pv.ctypes.data_as(_ffi.u32p),
                                                      C.byref(proof), C.byref(plen))
        else:
            rc = self._l.zkir_b200_prove_writelog(self._h, C.byref(params), pcs.ctypes.data, ins.ctypes.data, wlog.ctypes.data, n, int(wl["final_pc"]),
                                                  int(wl["entry_point"]), int(wl["exit_code"]), int(wl["halt_kind"]), log_n, pv.ctypes.data_as(_ffi.u32p),
                                                  C.byref(proof), C.byref(plen))
        self._check(rc)
        out = C.string_at(proof, plen.value)
        self._l.zkir_b200_free_proof(proof)
        return out, pv

    def prove_program(self, program, inputs, cfg):
        """Program -> Proof in one call (zkir_b200_prove_program): the interpreter records the write log into pinned memory and the
        log is uploaded chunk by chunk while it runs; for a program that needs the full profile the interpreter also records the memory
        log (28 B/cycle in all) and the device expands the wide table.  Returns (proof bytes, public values, cycles, log_n)."""
        params = cfg.params(profile_width(program_profile(program)))
        code = np.ascontiguousarray(program.code, dtype=np.uint32)
        data = np.ascontiguousarray(list(program.data) or [0], dtype=np.uint8)
        inp = np.ascontiguousarray([v & (2**64 - 1) for v in inputs] or [0], dtype=np.uint64)
        pv = np.zeros(NUM_PUBLIC, dtype=np.uint32)
        cycles, log_n = C.c_uint64(), C.c_uint32()
        proof, plen = C.c_void_p(), C.c_size_t()
        self._l.zkir_vm_enable_poseidon2(int(getattr(cfg, "enable_poseidon2_syscall", False)))
        rc = self._l.zkir_b200_prove_program(self._h, C.byref(params), code.ctypes.data, int(code.shape[0]), data.ctypes.data, len(program.data),
                                             program.entry_point, inp.ctypes.data, len(inputs), cfg.max_cycles, pv.ctypes.data_as(_ffi.u32p),
                                             C.byref(cycles), C.byref(log_n), C.byref(proof), C.byref(plen))
        self._check(rc)
        out = C.string_at(proof, plen.value)
        self._l.zkir_b200_free_proof(proof)
        return out, pv, int(cycles.value), int(log_n.value)

    def expand_writelog(self, wl, log_n, d_cols):
        if wl.get("program") is not None:
            self.set_program(wl["program"])
        n = int(wl["pcs"].shape[0])
        pcs, ins, wlog = (np.ascontiguousarray(wl[k]) for k in ("pcs", "instrs", "wlog"))
        self._check(self._l.zkir_b200_expand_writelog(self._h, pcs.ctypes.data, ins.ctypes.data, wlog.ctypes.data, n, int(wl["final_pc"]), log_n, d_cols))

    def expand_rows(self, rows, log_n, d_cols, profile="core"):
        if rows.get("program") is not None:
            self.set_program(rows["program"])
        n = int(rows["pcs"].shape[0])
        pcs, ins, regs = (np.ascontiguousarray(rows[k]) for k in ("pcs", "instrs", "regs"))
        fr = np.ascontiguousarray(rows["final_regs"], dtype=np.uint64)
        fn = self._l.zkir_b200_expand_rows_full if profile == "full" else self._l.zkir_b200_expand_rows
        self._check(fn(self._h, pcs.ctypes.data, ins.ctypes.data, regs.ctypes.data, n, fr.ctypes.data_as(_ffi.u64p),
                                                  int(rows["final_pc"]), log_n, d_cols))

    # -- per-kernel entry points (device pointers)
    def ntt(self, d_cols, n_cols, log_n, inverse=False, coset_shift=0):
        self._check(self._l.zkir_b200_ntt(self._h, d_cols, n_cols, log_n, int(inverse), coset_shift))

    def lde(self, d_in, d_out, n_cols, log_n, log_blowup):
        self._check(self._l.zkir_b200_lde(self._h, d_in, d_out, n_cols, log_n, log_blowup))

    def poseidon2_permute(self, d_states, n):
        self._check(self._l.zkir_b200_poseidon2_permute(self._h, d_states, n))

    def merkle_commit(self, d_matrix, n_cols, log_rows, d_tree):
        root = np.zeros(8, dtype=np.uint32)
        self._check(self._l.zkir_b200_merkle_commit(self._h, d_matrix, n_cols, log_rows, d_tree, root.ctypes.data_as(_ffi.u32p)))
        return root

    def quotient(self, cfg, d_lde, d_publde, log_n, public_values, lookup, alpha, d_q):
        params = cfg.params()
        pv = np.ascontiguousarray(public_values, dtype=np.uint32)
        al = np.ascontiguousarray(alpha, dtype=np.uint32)
        lk = np.ascontiguousarray(lookup, dtype=np.uint32)
        self._check(self._l.zkir_b200_quotient(self._h, C.byref(params), d_lde, d_publde, log_n, pv.ctypes.data_as(_ffi.u32p), lk.ctypes.data_as(_ffi.u32p),
                                               al.ctypes.data_as(_ffi.u32p), d_q))

    def aux_columns(self, d_trace, log_n, lookup, d_aux):
        lk = np.ascontiguousarray(lookup, dtype=np.uint32)
        self._check(self._l.zkir_b200_aux_columns(self._h, d_trace, log_n, lk.ctypes.data_as(_ffi.u32p), d_aux))

    def fri_fold(self, d_in, d_out, log_n, shift, beta):
        b = np.ascontiguousarray(beta, dtype=np.uint32)
        self._check(self._l.zkir_b200_fri_fold(self._h, d_in, d_out, log_n, shift, b.ctypes.data_as(_ffi.u32p)))


class PinnedBuffer:
    """array in page-locked host memory (zkir_b200_alloc_pinned) the interpreter's recorder / packer writes into."""

    def __init__(self, shape, dtype=np.uint32):
        self._l = _ffi.lib()
        dt = np.dtype(dtype)
        count = int(np.prod(shape))
        self._p = self._l.zkir_b200_alloc_pinned(max(count * dt.itemsize, 16))
        if not self._p:
            raise RuntimeError_(_ffi.ERR_OOM, "zkir_b200_alloc_pinned failed (needs a CUDA device)")
        raw = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), shape=(max(count * dt.itemsize, 16),))
        self.array = raw[:count * dt.itemsize].view(dt).reshape(shape)

    def close(self):
        if self._p:
            self.array = None
            self._l.zkir_b200_free_pinned(self._p)
            self._p = None

    __del__ = close


_default_ctx = {}


def _ctx(device):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def prove(program, inputs=(), cfg=None):
    """Program -> Proof: run the interpreter with the register write log recorded as it executes, rebuild + convert the rows on
    the GPU, prove.  The drop-in the north star describes for `zkir_runtime::prove()` (absent upstream: lib.rs:29-62)."""
    cfg = cfg or ProverConfig()
    ctx = _ctx(cfg.device)
    pb, pv, cycles, log_n = ctx.prove_program(program, list(inputs), cfg)   # either profile: zkir_b200_prove_program takes the width the program needs
    # the statement's public I/O transcript and outputs: one plain interpreter run without any recording (323 M cycles/s)
    res = VM(program, list(inputs), VMConfig(max_cycles=cfg.max_cycles, enable_poseidon2_syscall=cfg.enable_poseidon2_syscall)).run()
    return Proof(pb, pv, log_n, cycles, res.outputs, ctx.stage_ms(), program, res.io)


def verify(proof, cfg=None, public_values=None, program=None, io=None):
    """CPU verification through the C ABI; returns (ok, reason).  The statement is (program, public values, public I/O transcript):
    `program` = Program / code words, `io` = uint32 [n, 4] as `ExecutionResult.io`; passing an `ExecutionResult` as `program`
    supplies both, and a `Proof` returned by prove() carries them."""
    cfg = cfg or ProverConfig()
    l = _ffi.lib()
    pb = proof.bytes_ if isinstance(proof, Proof) else bytes(proof)
    pv = public_values if public_values is not None else (proof.public_values if isinstance(proof, Proof) else None)
    if program is None and isinstance(proof, Proof):
        program, io = proof.program, (proof.io if io is None else io)
    if isinstance(program, ExecutionResult):
        program, io = program.program, (program.io if io is None else io)
    if program is None:
        return False, "verify() needs the program the proof is about"
    code = np.ascontiguousarray(getattr(program, "code", program), dtype=np.uint32)
    ev = np.ascontiguousarray(io if io is not None else np.zeros((0, 4)), dtype=np.uint32).reshape(-1, 4)
    params = cfg.params(int.from_bytes(pb[12:16], "little") if len(pb) >= 16 else WIDTH)   # the profile is the proof header's width
    buf = C.create_string_buffer(pb, len(pb))
    pvp = np.ascontiguousarray(pv, dtype=np.uint32).ctypes.data_as(_ffi.u32p) if pv is not None else None
    rc = l.zkir_b200_verify(C.byref(params), C.cast(buf, C.c_void_p), len(pb), pvp, code.ctypes.data, int(code.shape[0]), ev.ctypes.data, int(ev.shape[0]))
    return rc == 0, ("" if rc == 0 else l.zkir_b200_last_error(None).decode())
