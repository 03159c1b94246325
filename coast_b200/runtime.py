"""ctypes mirror of include/coast_rt.h.

Error behaviour mirrors the C ABI: every call that fails raises :class:`CoastError` carrying
``coast_last_error()``.  There is no CPU fallback anywhere in this module -- without the CUDA
driver ``Runtime()`` raises (COAST_ERR_NO_DRIVER), loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

HERE = os.path.dirname(os.path.abspath(__file__))

K_CRC16, K_SHA256, K_AES128, K_MM_U32, K_GEMM_TF32, K_QSORT, K_CHSTONE_SHA, K_CHSTONE_AES = range(8)
F_COUNT_ERRORS, F_COUNT_SYNCS, F_NO_MEM_REPLICATION = 0x1, 0x2, 0x4
F_INTERLEAVE, F_SEGMENT, F_VERBOSE, F_MAJORITY_VOTER = 0x8, 0x10, 0x20, 0x100
F_STORE_DATA_SYNC, F_NO_STORE_DATA_SYNC, F_NO_LOAD_SYNC, F_NO_STORE_ADDR_SYNC = 0x200, 0x400, 0x800, 0x1000
PLAN_NONE, PLAN_BERNOULLI, PLAN_TABLE = 0, 1, 2
AES_DECRYPT, AES_KEY_PER_UNIT, AES_KEY_WRITEBACK = 1, 2, 4
NO_FAULT_UNIT = 0xFFFFFFFFFFFFFFFF
ERR_NO_DRIVER, ERR_NOT_INIT, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_BUSY = -100001, -100002, -100003, -100004, -100005

OUT_BYTES = {K_CRC16: 2, K_SHA256: 32, K_AES128: 16, K_MM_U32: 4, K_GEMM_TF32: 4, K_CHSTONE_SHA: 20, K_CHSTONE_AES: 64}


def out_bytes(kernel: int, unit_bytes: int = 0) -> int:
    return unit_bytes if kernel == K_QSORT else OUT_BYTES[kernel]


class CoastError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"coast_rt error {code}: {msg}")
        self.code = code


class _Plan(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32), ("threshold", C.c_uint32),
                ("d_table", C.c_void_p)]


class LaunchDesc(C.Structure):
    _fields_ = [("kernel", C.c_uint32), ("num_clones", C.c_uint32), ("flags", C.c_uint32), ("mode", C.c_uint32),
                ("n_units", C.c_uint64), ("unit_base", C.c_uint64),
                ("unit_bytes", C.c_uint32), ("M", C.c_uint32), ("N", C.c_uint32), ("K", C.c_uint32),
                ("d_in", C.c_void_p), ("d_out", C.c_void_p), ("d_aux", C.c_void_p),
                ("key", C.c_uint8 * 16), ("plan", C.POINTER(_Plan)), ("d_status", C.c_void_p)]


class _Stats(C.Structure):
    _fields_ = [("errors_corrected", C.c_uint64), ("dwc_detected", C.c_uint64), ("syncs", C.c_uint64),
                ("injected", C.c_uint64), ("first_fault_unit", C.c_uint64)]


@dataclass
class Stats:
    errors_corrected: int = 0
    dwc_detected: int = 0
    syncs: int = 0
    injected: int = 0
    first_fault_unit: int = NO_FAULT_UNIT

    def as_dict(self):
        return dict(errors_corrected=self.errors_corrected, dwc_detected=self.dwc_detected, syncs=self.syncs,
                    injected=self.injected, first_fault_unit=self.first_fault_unit)


@dataclass
class FaultPlan:
    """On-device single-bit-flip plan (include/coast_rt.h, "Fault plan")."""
    mode: int = PLAN_NONE
    seed: int = 0
    p: float = 0.0
    threshold: int | None = None
    table: object = None  # torch.uint32/int32 CUDA tensor, one entry per local unit (PLAN_TABLE)

    def to_c(self) -> _Plan:
        pl = _Plan()
        pl.mode = self.mode
        pl.seed_lo = self.seed & 0xFFFFFFFF
        pl.seed_hi = (self.seed >> 32) & 0xFFFFFFFF
        thr = self.threshold if self.threshold is not None else min(int(self.p * 2 ** 32), 0xFFFFFFFF)
        pl.threshold = thr
        pl.d_table = self.table.data_ptr() if self.table is not None else None
        return pl


def fault_entry(replica: int, site: int, bit: int) -> int:
    return 0x80000000 | ((replica & 3) << 29) | ((site & 0xFFFFFF) << 5) | (bit & 31)


def lib_path() -> str:
    return os.path.join(HERE, "libcoast_rt.so")


_LIB = None

EXPORTS = [
    "coast_init", "coast_numa_node", "coast_shutdown", "coast_last_error", "coast_version", "coast_parse_opt_passes", "coast_flags_honoured", "coast_launch",
    "coast_sync", "coast_sync_noabort", "coast_stats_snapshot", "coast_stats_reset", "coast_counters_export", "coast_counters_attach",
    "coast_counters_detach", "coast_sm_count", "coast_clock_probe", "coast_fault_sites",
    "coast_fault_site_bits", "coast_out_bytes_per_unit", "coast_out_bytes", "coast_votes_per_unit", "coast_malloc", "coast_free",
    "coast_memcpy_h2d", "coast_memcpy_d2h", "coast_memset", "coast_host_alloc", "coast_host_free",
    "coast_stream_create", "coast_stream_destroy", "coast_stream_sync", "coast_fill_philox", "coast_run_host",
    "coast_run_host_noabort", "coast_last_host_path",
    "coast_set_opt_passes", "coast_xmr_crc16", "coast_xmr_sha256_hash", "coast_xmr_aes_enc_dec",
    "coast_xmr_matrix_multiply_u32", "coast_xmr_chstone_sha_stream", "coast_xmr_chstone_aes", "TMR_ERROR_CNT", "__SYNC_COUNT", "FAULT_DETECTED_DWC",
]


def load_library():
    """dlopen the in-tree libcoast_rt.so; raises if it has not been built (no silent fallback)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise CoastError(ERR_NOT_INIT, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                           "(or make -C coast_b200/csrc); there is no CPU fallback")
        L = C.CDLL(path)
        L.coast_last_error.restype = C.c_char_p
        L.coast_version.restype = C.c_char_p
        L.coast_last_host_path.restype = C.c_char_p
        L.coast_init.argtypes = [C.c_int]
        L.coast_parse_opt_passes.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.coast_set_opt_passes.argtypes = [C.c_char_p]
        L.coast_flags_honoured.argtypes = [C.c_uint32] * 3
        L.coast_flags_honoured.restype = C.c_uint32
        L.coast_launch.argtypes = [C.POINTER(LaunchDesc), C.c_void_p]
        L.coast_run_host.argtypes = [C.POINTER(LaunchDesc), C.POINTER(_Stats)]
        L.coast_run_host_noabort.argtypes = [C.POINTER(LaunchDesc), C.POINTER(_Stats)]
        L.coast_sync.argtypes = [C.c_void_p, C.POINTER(_Stats)]
        L.coast_sync_noabort.argtypes = [C.c_void_p, C.POINTER(_Stats)]
        L.coast_stats_snapshot.argtypes = [C.c_void_p, C.c_void_p]
        L.coast_stats_reset.argtypes = [C.c_void_p]
        L.coast_counters_export.argtypes = [C.c_void_p]
        L.coast_clock_probe.argtypes = [C.c_void_p, C.c_void_p]
        L.coast_counters_attach.argtypes = [C.c_void_p]
        L.coast_fault_sites.argtypes = [C.c_uint32] * 3
        L.coast_fault_sites.restype = C.c_uint32
        L.coast_fault_site_bits.argtypes = [C.c_uint32] * 4
        L.coast_fault_site_bits.restype = C.c_uint32
        L.coast_out_bytes_per_unit.argtypes = [C.c_uint32]
        L.coast_out_bytes_per_unit.restype = C.c_uint32
        L.coast_votes_per_unit.argtypes = [C.c_uint32]
        L.coast_votes_per_unit.restype = C.c_uint32
        L.coast_fill_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.coast_xmr_crc16.argtypes = [C.c_char_p, C.c_ubyte]
        L.coast_xmr_crc16.restype = C.c_ushort
        _LIB = L
    return _LIB


def parse_opt_passes(opt_passes: str) -> tuple[int, int]:
    """OPT_PASSES string of a reference test Makefile -> (num_clones, flags)."""
    L = load_library()
    nc, fl = C.c_uint32(), C.c_uint32()
    rc = L.coast_parse_opt_passes(opt_passes.encode(), C.byref(nc), C.byref(fl))
    if rc:
        raise CoastError(rc, L.coast_last_error().decode())
    return nc.value, fl.value


class Runtime:
    """One process <-> one GPU.  Device memory and streams come from torch (plumbing)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.device = device
        self._check(self.L.coast_init(device))
        import torch  # after coast_init so a missing driver is reported by OUR library, loudly
        self.torch = torch
        torch.cuda.set_device(device)

    def _check(self, rc: int):
        if rc != 0:
            raise CoastError(rc, self.L.coast_last_error().decode())

    # -- low level -------------------------------------------------------------------------
    def stream_handle(self, stream=None) -> int:
        s = stream if stream is not None else self.torch.cuda.current_stream()
        return s.cuda_stream

    def make_desc(self, kernel, num_clones, d_in, d_out, n_units, *, flags=0, mode=0, unit_bytes=0, M=0, N=0, K=0,
                  d_aux=None, key: bytes | None = None, plan: FaultPlan | None = None, unit_base=0, d_status=None):
        d = LaunchDesc()
        d.kernel, d.num_clones, d.flags, d.mode = kernel, num_clones, flags, mode
        d.n_units, d.unit_base, d.unit_bytes = n_units, unit_base, unit_bytes
        d.M, d.N, d.K = M, N, K
        d.d_in = d_in.data_ptr() if hasattr(d_in, "data_ptr") else d_in
        d.d_out = d_out.data_ptr() if hasattr(d_out, "data_ptr") else d_out
        if d_aux is not None:
            d.d_aux = d_aux.data_ptr() if hasattr(d_aux, "data_ptr") else d_aux
        if key is not None:
            d.key = (C.c_uint8 * 16)(*key)
        if d_status is not None:
            d.d_status = d_status.data_ptr() if hasattr(d_status, "data_ptr") else d_status
        keep = None
        if plan is not None and plan.mode != PLAN_NONE:
            keep = plan.to_c()
            d.plan = C.pointer(keep)
        d._keep = (keep, d_in, d_out, d_aux)
        return d

    def launch(self, desc: LaunchDesc, stream=None):
        self._check(self.L.coast_launch(C.byref(desc), self.stream_handle(stream)))

    def sync(self, stream=None, abort_on_dwc: bool = False) -> Stats:
        st = _Stats()
        fn = self.L.coast_sync if abort_on_dwc else self.L.coast_sync_noabort
        self._check(fn(self.stream_handle(stream), C.byref(st)))
        return Stats(st.errors_corrected, st.dwc_detected, st.syncs, st.injected, st.first_fault_unit)

    def stats_snapshot(self, d_out, stream=None):
        self._check(self.L.coast_stats_snapshot(self.stream_handle(stream), d_out.data_ptr()))

    def sm_count(self) -> int:
        return int(self.L.coast_sm_count())

    def clock_probe(self, d_out, stream=None):
        """d_out: int64 CUDA tensor of 2 * sm_count() elements -> {clock64, globaltimer ns} per SM"""
        self._check(self.L.coast_clock_probe(C.c_void_p(d_out.data_ptr()), self.stream_handle(stream)))

    # multi-GPU fold of the counters over NVLink peer memory (include/coast_rt.h: coast_counters_export / _attach / _detach)
    def counters_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.L.coast_counters_export(buf))
        return buf.raw

    def counters_attach(self, handle: bytes):
        assert len(handle) == 64
        self._check(self.L.coast_counters_attach(C.create_string_buffer(handle, 64)))

    def counters_detach(self):
        self._check(self.L.coast_counters_detach())

    def stats_reset(self, stream=None):
        self._check(self.L.coast_stats_reset(self.stream_handle(stream)))

    def fill_philox(self, dst, seed: int, word_base: int = 0, stream=None):
        """dst: any CUDA tensor whose byte size is a multiple of 4."""
        nbytes = dst.numel() * dst.element_size()
        assert nbytes % 4 == 0
        self._check(self.L.coast_fill_philox(dst.data_ptr(), nbytes // 4, word_base, seed, self.stream_handle(stream)))

    def fault_sites(self, kernel, unit_bytes=0, K=0) -> int:
        return int(self.L.coast_fault_sites(kernel, unit_bytes, K))

    def fault_site_bits(self, kernel, unit_bytes, K, site) -> int:
        return int(self.L.coast_fault_site_bits(kernel, unit_bytes, K, site))

    @property
    def numa_node(self) -> int:
        return int(self.L.coast_numa_node())

    @property
    def last_host_path(self) -> str:
        return self.L.coast_last_host_path().decode()

    @property
    def tmr_error_cnt(self) -> int:
        return C.c_uint32.in_dll(self.L, "TMR_ERROR_CNT").value

    @property
    def sync_count(self) -> int:
        return C.c_uint64.in_dll(self.L, "__SYNC_COUNT").value

    # -- convenience: device tensors in, device tensor + Stats out ----------------------------
    def run(self, kernel, num_clones, inp, n_units, *, flags=0, mode=0, unit_bytes=0, M=0, N=0, K=0, aux=None,
            key: bytes | None = None, plan: FaultPlan | None = None, unit_base=0, out=None, stream=None, status=None):
        torch = self.torch
        if out is None:
            out = torch.empty(n_units * out_bytes(kernel, unit_bytes), dtype=torch.uint8, device=f"cuda:{self.device}")
        d = self.make_desc(kernel, num_clones, inp, out, n_units, flags=flags, mode=mode, unit_bytes=unit_bytes,
                           M=M, N=N, K=K, d_aux=aux, key=key, plan=plan, unit_base=unit_base, d_status=status)
        self.launch(d, stream)
        return out, self.sync(stream)

    # -- the reference-facing host call: HOST buffers, H2D + kernel + D2H inside ---------------
    def run_host(self, kernel, num_clones, h_in, h_out, n_units, *, flags=0, mode=0, unit_bytes=0, M=0, N=0, K=0,
                 h_aux=None, key: bytes | None = None, plan: FaultPlan | None = None, unit_base=0,
                 abort_on_dwc: bool = False) -> Stats:
        """h_in/h_out/h_aux: CPU torch tensors or numpy arrays (pinned memory makes the copies async)."""
        def ptr(x):
            if x is None:
                return None
            return x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data
        d = self.make_desc(kernel, num_clones, ptr(h_in), ptr(h_out), n_units, flags=flags, mode=mode,
                           unit_bytes=unit_bytes, M=M, N=N, K=K, d_aux=ptr(h_aux), key=key, plan=plan,
                           unit_base=unit_base)
        d._keep2 = (h_in, h_out, h_aux)
        st = _Stats()
        fn = self.L.coast_run_host if abort_on_dwc else self.L.coast_run_host_noabort
        self._check(fn(C.byref(d), C.byref(st)))
        return Stats(st.errors_corrected, st.dwc_detected, st.syncs, st.injected, st.first_fault_unit)
