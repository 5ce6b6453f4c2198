"""spectra_b200 — B200-native implicitly restarted Lanczos/Arnoldi eigensolver.

Python mirror of the reference's public interface for the one hot path:

    op   = SparseSymMatProd(A)            # MatOp/SparseSymMatProd.h   (lower triangle of A is read)
    eigs = SymEigsSolver(op, nev, ncv)    # SymEigsSolver.h / HermEigsBase.h
    eigs.init(); nconv = eigs.compute(SortRule.LargestAlge)
    eigs.info(), eigs.eigenvalues(), eigs.eigenvectors()

Every call goes through the C ABI of ``lib/libspectra_b200.so`` (``include/spectra_b200.h``):
hand-written sm_100a CUDA kernels driven by a C++ host loop.  There is no CPU fallback — if the
library is missing or no B200-class GPU is present the calls raise.  PyTorch is not used here at
all; ``torch.distributed`` is only the launcher's rendezvous in multi-GPU runs (see ``Comm``).
"""
from __future__ import annotations

import ctypes as C
import enum
import os

import numpy as np

__all__ = [
    "SortRule", "CompInfo", "SparseSymMatProd", "SparseGenMatProd", "SparseHermMatProd", "UserOp", "SymEigsSolver", "HermEigsSolver", "GenEigsSolver", "Comm", "Stats", "lib", "lib_path", "device_info",
    "set_profiling", "set_device", "dense",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
# SB200_LIB_SUFFIX selects an experimental build variant (developer knob, see _build.py); default: the product library
_LIB_PATH = os.path.join(_HERE, "lib", "libspectra_b200" + os.environ.get("SB200_LIB_SUFFIX", "") + ".so")
_lib = None


class SortRule(enum.IntEnum):
    """Util/SelectionRule.h:33-58"""
    LargestMagn = 0
    LargestReal = 1
    LargestImag = 2
    LargestAlge = 3
    SmallestMagn = 4
    SmallestReal = 5
    SmallestImag = 6
    SmallestAlge = 7
    BothEnds = 8


class CompInfo(enum.IntEnum):
    """Util/CompInfo.h:17-30"""
    Successful = 0
    NotComputed = 1
    NotConverging = 2
    NumericalIssue = 3


class Stats(C.Structure):
    _fields_ = [
        ("lanczos_steps", C.c_int64), ("reorth_passes", C.c_int64), ("restarts", C.c_int64), ("expand_calls", C.c_int64), ("kernel_launches", C.c_int64),
        ("spmv_launches", C.c_int64), ("panel_launches", C.c_int64), ("panel_cols", C.c_int64), ("compress_launches", C.c_int64), ("compress_cols", C.c_int64),
        ("ms_total", C.c_double), ("ms_spmv", C.c_double), ("ms_panel", C.c_double), ("ms_compress", C.c_double), ("ms_small", C.c_double),
        ("ms_comm", C.c_double), ("fused_dot_launches", C.c_int64), ("fused_dot_cols", C.c_int64), ("host_syncs", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib_path() -> str:
    return _LIB_PATH


def lib():
    """Loads the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: build it with `python -m spectra_b200._build` (nvcc, sm_100a). "
                               "spectra_b200 has no CPU fallback.")
        L = C.CDLL(_LIB_PATH)
        L.sb200_last_error.restype = C.c_char_p
        L.sb200_version.restype = C.c_char_p
        _lib = L
    return _lib


# status -> exception type, mirroring the reference's exceptions
class SpectraError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


class InvalidArgument(SpectraError, ValueError):
    """std::invalid_argument"""


class LogicError(SpectraError):
    """std::logic_error"""


class CudaError(SpectraError):
    pass


def _check(rc: int):
    if rc == 0:
        return
    msg = lib().sb200_last_error().decode(errors="replace")
    if rc == 1:
        raise InvalidArgument(rc, msg)
    if rc == 2:
        raise LogicError(rc, msg)
    if rc == 4:
        raise CudaError(rc, msg)
    raise SpectraError(rc, msg)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check_op(op, rc: int):
    """_check for calls that may run a Python operator callback: an exception raised inside the callback cannot cross
    the C frames, so it is parked on the operator and re-raised here once the C call has returned."""
    exc = getattr(op, "_exc", None)
    if exc is not None:
        op._exc = None
        raise exc
    _check(rc)


def device_info() -> dict:
    dev, sms, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    _check(lib().sb200_device_info(C.byref(dev), C.byref(sms), C.byref(maj), C.byref(mnr), C.byref(mem)))
    return dict(device=dev.value, sm_count=sms.value, cc=(maj.value, mnr.value), hbm_bytes=mem.value)


def set_device(device: int):
    """One process per GPU: select the CUDA device before any other call."""
    _check(lib().sb200_set_device(int(device)))


def set_profiling(level: int):
    _check(lib().sb200_set_profiling(int(level)))


class Comm:
    """Row-sharded multi-GPU communicator (one process per GPU).  `id128` comes from
    Comm.unique_id() on rank 0 and is distributed by the launcher (e.g. torch.distributed)."""

    def __init__(self, rank: int, nranks: int, id128: bytes | None):
        self.h = C.c_void_p()
        buf = C.create_string_buffer(id128, 128) if id128 is not None else None
        _check(lib().sb200_comm_create(int(rank), int(nranks), buf, C.byref(self.h)))
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().sb200_comm_unique_id(buf))
        return buf.raw

    def close(self):
        if self.h:
            lib().sb200_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _result_buffer(y_out, count, dtype):
    """(array handed to the C ABI, user array to copy back into or None).  A caller-supplied y_out is written in place only when it
    is a C-contiguous array of exactly the element type and length the library writes; anything else (float32, strided, a list) gets
    the result through a temporary so that a wrong buffer can never be overrun."""
    if y_out is None:
        return np.empty(count, dtype=dtype), None
    if not isinstance(y_out, np.ndarray) or y_out.size != count:
        raise InvalidArgument(1, "y_out has the wrong length")
    if not y_out.flags.writeable:
        raise InvalidArgument(1, "y_out is read-only")
    if y_out.dtype == np.dtype(dtype) and y_out.flags.c_contiguous:
        return y_out, None
    return np.empty(count, dtype=dtype), y_out



class _SparseOp:
    """Device-resident CSR operator.  `mat` is a scipy.sparse CSC/CSR matrix or a tuple
    (n, outer, inner, values, 'col'|'row') with Eigen's compressed layout."""
    _mode = 0

    _shift_solve = False

    def __init__(self, mat, comm: Comm | None = None, uplo: str = "lower"):
        n, outer, inner, vals, order = self._unpack(mat)
        self.n = int(n)
        self._keep = (outer, inner, vals)  # like Eigen::Ref: the wrapper refers to the user's arrays
        self._order = order
        mode = self._mode
        if mode != 0:
            mode = 1 if uplo == "lower" else 2
        outer64 = 1 if outer.dtype == np.int64 else 0
        self.h = C.c_void_p()
        self.comm = comm
        if self._shift_solve:
            _check(lib().sb200_op_create_shift_solve(C.c_int64(self.n), _p(outer), outer64, _p(inner), _p(vals), 0 if order == "col" else 1, mode, C.byref(self.h)))
        else:
            _check(lib().sb200_op_create_sparse(C.c_int64(self.n), _p(outer), outer64, _p(inner), _p(vals), 0 if order == "col" else 1, mode,
                                                comm.h if comm is not None else None, C.byref(self.h)))
        r0, nr = C.c_int64(), C.c_int64()
        _check(lib().sb200_op_local_rows(self.h, C.byref(r0), C.byref(nr)))
        self.row0, self.nrows_local = r0.value, nr.value

    @staticmethod
    def _unpack(mat):
        if isinstance(mat, tuple):
            n, outer, inner, vals, order = mat
        else:
            import scipy.sparse as sp

            if sp.isspmatrix_csr(mat):
                order = "row"
            else:
                mat = sp.csc_matrix(mat)
                order = "col"
            if mat.shape[0] != mat.shape[1]:
                raise InvalidArgument(1, "matrix must be square")
            n, outer, inner, vals = mat.shape[0], mat.indptr, mat.indices, mat.data
        outer = np.ascontiguousarray(outer)
        if outer.dtype not in (np.int32, np.int64):
            outer = outer.astype(np.int64)
        inner = np.ascontiguousarray(inner, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        return n, outer, inner, vals, order

    @classmethod
    def from_csr_slab(cls, n, row0, rowptr_local, col, vals, comm: Comm | None = None):
        """Pre-partitioned rows [row0, row0+len(rowptr_local)-1) of a full CSR (general semantics)."""
        self = cls.__new__(cls)
        self.n = int(n)
        rowptr_local = np.ascontiguousarray(rowptr_local, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self._keep = (rowptr_local, col, vals)
        self._order = "row"
        self.h = C.c_void_p()
        self.comm = comm
        nrows = len(rowptr_local) - 1
        _check(lib().sb200_op_create_csr_slab(C.c_int64(self.n), C.c_int64(row0), C.c_int64(nrows), _p(rowptr_local), _p(col), _p(vals),
                                              comm.h if comm is not None else None, C.byref(self.h)))
        self.row0, self.nrows_local = int(row0), int(nrows)
        return self

    def rows(self):
        return self.n

    def cols(self):
        return self.n

    @property
    def nnz(self):
        v = C.c_int64()
        _check(lib().sb200_op_nnz(self.h, C.byref(v)))
        return v.value

    def perform_op(self, x_in: np.ndarray, y_out: np.ndarray | None = None) -> np.ndarray:
        """y_out = A * x_in with host arrays (SparseSymMatProd.h:83-88)."""
        x = np.ascontiguousarray(x_in, dtype=np.float64)
        if x.shape != (self.n,):
            raise InvalidArgument(1, "x has the wrong length")
        y, user = _result_buffer(y_out, self.nrows_local, np.float64)
        _check(lib().sb200_op_perform_op(self.h, _p(x), _p(y)))
        if user is not None:
            user.reshape(-1)[:] = y
            return user
        return y

    def __matmul__(self, X):
        """operator*(Matrix) (SparseSymMatProd.h:93-96)"""
        X = np.asfortranarray(X, dtype=np.float64)
        if X.ndim == 1:
            return self.perform_op(X)
        Y = np.empty((self.nrows_local, X.shape[1]), order="F")
        _check(lib().sb200_op_apply_matrix(self.h, _p(X), C.c_int64(X.shape[1]), _p(Y)))
        return Y

    def __call__(self, i, j):
        """operator()(i, j) (SparseSymMatProd.h:101-104): the STORED coefficient of the user's matrix."""
        outer, inner, vals = self._keep
        o, k = (j, i) if self._order == "col" else (i, j)
        seg = inner[outer[o]:outer[o + 1]]
        hit = np.nonzero(seg == k)[0]
        return float(vals[outer[o] + hit[0]]) if len(hit) else 0.0

    def spmv_device_time(self, repeat: int = 10) -> float:
        """Average device time (ms) of one CSR SpMV launch on a device-resident vector (benchmark hook)."""
        ms = C.c_float()
        _check(lib().sb200_op_spmv_device(self.h, None, None, 3, C.byref(ms)))  # warm-up
        _check(lib().sb200_op_spmv_device(self.h, None, None, int(repeat), C.byref(ms)))
        return ms.value / repeat

    def spmv_layout(self) -> dict:
        """Device layout: format ('csr' | 'sell'), column blocks, stored entries (incl. padding of the sliced layout)."""
        fmt, nb, stored = C.c_int(), C.c_int(), C.c_int64()
        _check(lib().sb200_op_layout_info(self.h, C.byref(fmt), C.byref(nb), C.byref(stored)))
        return dict(format="sell" if fmt.value == 1 else "csr", col_blocks=nb.value, stored_entries=stored.value)

    def peer_mode(self) -> bool:
        """True when a row-sharded operator exchanges operand and dot products through NVLink-mapped peer memory (False: NCCL collectives)."""
        v = C.c_int()
        _check(lib().sb200_op_peer_mode(self.h, C.byref(v)))
        return bool(v.value)

    def close(self):
        if getattr(self, "h", None):
            lib().sb200_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bench_gather(n: int, gathers: int, repeat: int = 5) -> dict:
    """Roofline microbenchmark (tools/gather_roof.py): time of `gathers` uniformly random 8-byte loads from n doubles."""
    ms, chk = C.c_float(), C.c_double()
    _check(lib().sb200_bench_gather(C.c_int64(int(n)), C.c_int64(int(gathers)), int(repeat), C.byref(ms), C.byref(chk)))
    return dict(ms=ms.value, checksum=chk.value)


def bench_stream_gather(n: int, gathers: int, band: bool = False, repeat: int = 5) -> dict:
    """Roofline microbenchmark: streamed (index, coefficient) pairs + dependent gathers, the floor of any SpMV on this access pattern."""
    ms, chk = C.c_float(), C.c_double()
    _check(lib().sb200_bench_stream_gather(C.c_int64(int(n)), C.c_int64(int(gathers)), int(bool(band)), int(repeat), C.byref(ms), C.byref(chk)))
    return dict(ms=ms.value, checksum=chk.value)


_MATVEC_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


class UserOp:
    """User-defined operator — the reference's OpType concept (SymEigsSolver.h:99-114): `op` is either a callable
    y = op(x) on numpy vectors, or an object with rows() and perform_op(x_in, y_out).  The Krylov iteration still
    runs on the GPU; every matrix operation round-trips one vector through pinned host memory."""

    def __init__(self, op, n: int | None = None, complex_scalar: bool = False):
        """complex_scalar=True: the operator works on complex128 vectors (Scalar = std::complex<double>, for HermEigsSolver)."""
        dt = np.complex128 if complex_scalar else np.float64
        self._dtype = dt
        if hasattr(op, "perform_op"):
            n = int(op.rows()) if n is None else int(n)

            def apply(x, y):
                op.perform_op(x, y)
        else:
            if n is None:
                raise InvalidArgument(1, "n is required for a callable operator")
            n = int(n)

            def apply(x, y):
                y[:] = op(x)

        self.n = n
        self.row0, self.nrows_local = 0, n
        self.comm = None
        self._exc = None
        self.user = op

        cw = 2 if complex_scalar else 1

        def tramp(xp, yp, _):
            try:
                # complex vectors cross the boundary as interleaved (re, im) doubles: view them as complex128 without copying
                apply(np.ctypeslib.as_array(xp, shape=(cw * n,)).view(dt), np.ctypeslib.as_array(yp, shape=(cw * n,)).view(dt))
            except BaseException as e:  # noqa: BLE001 - re-raised on the Python side after the C call returns
                self._exc = e

        self._cb = _MATVEC_FN(tramp)  # must outlive the operator handle
        self.h = C.c_void_p()
        create = lib().sb200_op_create_callback_z if complex_scalar else lib().sb200_op_create_callback
        _check(create(C.c_int64(n), self._cb, None, C.byref(self.h)))

    def rows(self):
        return self.n

    def cols(self):
        return self.n

    def set_shift(self, sigma):
        """Forwarded to the user's shift-solve operator (SymEigsShiftSolver.h:194 calls op.set_shift(sigma))."""
        if hasattr(self.user, "set_shift"):
            self.user.set_shift(sigma)

    def perform_op(self, x_in, y_out=None):
        x = np.ascontiguousarray(x_in, dtype=self._dtype)
        y, user = _result_buffer(y_out, self.n, self._dtype)
        _check_op(self, lib().sb200_op_perform_op(self.h, _p(x), _p(y)))
        if user is not None:
            user.reshape(-1)[:] = y
            return user
        return y

    def close(self):
        if getattr(self, "h", None):
            lib().sb200_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SparseSymMatProd(_SparseOp):
    """MatOp/SparseSymMatProd.h — reads only the `uplo` triangle of the stored matrix."""
    _mode = 1

    def __init__(self, mat, uplo: str = "lower", comm: Comm | None = None):
        super().__init__(mat, comm=comm, uplo=uplo)


class SparseGenMatProd(_SparseOp):
    """MatOp/SparseGenMatProd.h"""
    _mode = 0

    def __init__(self, mat, comm: Comm | None = None):
        super().__init__(mat, comm=comm)


class SparseHermMatProd:
    """MatOp/SparseHermMatProd.h with Scalar = std::complex<double>: y = A.selfadjointView<Uplo>() * x for a complex sparse
    matrix of which only one triangle is read (mirrored conjugated; the diagonal is taken as real).  `mat` is a complex
    scipy.sparse CSC/CSR matrix or a tuple (n, outer, inner, values[complex128], 'col'|'row').  Vectors are complex128.
    Experimental in round 1: verified on the kernel-logic emulator, not yet on a device."""

    def __init__(self, mat, uplo: str = "lower"):
        if isinstance(mat, tuple):
            n, outer, inner, vals, order = mat
        else:
            import scipy.sparse as sp

            if sp.isspmatrix_csr(mat):
                order = "row"
            else:
                mat = sp.csc_matrix(mat)
                order = "col"
            if mat.shape[0] != mat.shape[1]:
                raise InvalidArgument(1, "matrix must be square")
            n, outer, inner, vals = mat.shape[0], mat.indptr, mat.indices, mat.data
        outer = np.ascontiguousarray(outer)
        if outer.dtype not in (np.int32, np.int64):
            outer = outer.astype(np.int64)
        inner = np.ascontiguousarray(inner, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.complex128)
        self.n = int(n)
        self._keep = (outer, inner, vals)
        self._order = order
        self.comm = None
        self.row0, self.nrows_local = 0, self.n
        mode = {"lower": 3, "upper": 4, "general": 0}[uplo]
        self.h = C.c_void_p()
        _check(lib().sb200_op_create_sparse_herm(C.c_int64(self.n), _p(outer), 1 if outer.dtype == np.int64 else 0, _p(inner), _p(vals),
                                                 0 if order == "col" else 1, mode, C.byref(self.h)))

    def rows(self):
        return self.n

    def cols(self):
        return self.n

    @property
    def nnz(self):
        v = C.c_int64()
        _check(lib().sb200_op_nnz(self.h, C.byref(v)))
        return v.value

    def perform_op(self, x_in: np.ndarray, y_out: np.ndarray | None = None) -> np.ndarray:
        x = np.ascontiguousarray(x_in, dtype=np.complex128)
        if x.shape != (self.n,):
            raise InvalidArgument(1, "x has the wrong length")
        y, user = _result_buffer(y_out, self.n, np.complex128)
        _check(lib().sb200_op_perform_op(self.h, _p(x), _p(y)))
        if user is not None:
            user.reshape(-1)[:] = y
            return user
        return y

    def __matmul__(self, X):
        X = np.asfortranarray(X, dtype=np.complex128)
        if X.ndim == 1:
            return self.perform_op(X)
        Y = np.empty((self.n, X.shape[1]), dtype=np.complex128, order="F")
        _check(lib().sb200_op_apply_matrix(self.h, _p(X), C.c_int64(X.shape[1]), _p(Y)))
        return Y

    def close(self):
        if getattr(self, "h", None):
            lib().sb200_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SparseSymShiftSolve(_SparseOp):
    """MatOp/SparseSymShiftSolve.h — perform_op computes y = (A - sigma I)^{-1} x after set_shift(sigma).  Device
    implementation: block cyclic reduction for half-bandwidth <= 32, sequential block elimination with grid-wide block kernels for wider
    bands / mesh-like patterns (as long as the 3 n b doubles of block factors fit in device memory), explicit inverse for n <= 2048."""
    _mode = 1
    _shift_solve = True

    def __init__(self, mat, uplo: str = "lower"):
        super().__init__(mat, comm=None, uplo=uplo)

    def set_shift(self, sigma: float):
        _check(lib().sb200_op_set_shift(self.h, C.c_double(sigma)))

    def set_refine(self, steps: int):
        """0 / 1: fixed number of refinement steps per solve; negative: set_shift() decides (the default)."""
        _check(lib().sb200_op_shift_solve_refine(self.h, int(steps)))

    def status(self) -> dict:
        """Outcome of the last set_shift(): refinement steps in use and the relative residuals of its verification solve."""
        steps, rel, rel0 = C.c_int(), C.c_double(), C.c_double()
        _check(lib().sb200_op_shift_solve_status(self.h, C.byref(steps), C.byref(rel), C.byref(rel0)))
        return dict(refine_steps=steps.value, verify_residual=rel.value, unrefined_residual=rel0.value)

    def layout(self) -> dict:
        bw, blk, lev, rows = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        _check(lib().sb200_op_shift_solve_info(self.h, C.byref(bw), C.byref(blk), C.byref(rows), C.byref(lev)))
        return dict(half_bandwidth=bw.value, block=blk.value, block_rows=rows.value, levels=lev.value)

    def solve_device_time(self, repeat: int = 10) -> float:
        """CUDA-event time (ms) of one device-resident solve."""
        return self.spmv_device_time(repeat)


class SymEigsSolver:
    """SymEigsSolver.h:133-160 / HermEigsBase.h — same call sequence as the reference."""

    def __init__(self, op: _SparseOp, nev: int, ncv: int):
        self.op = op
        self.nev, self.ncv = int(nev), int(ncv)
        self.h = C.c_void_p()
        self._create(op, nev, ncv)

    def _create(self, op, nev, ncv):
        _check(lib().sb200_sym_create(op.h, C.c_int64(nev), C.c_int64(ncv), C.byref(self.h)))

    def init(self, init_resid: np.ndarray | None = None):
        r = np.ascontiguousarray(init_resid, dtype=np.float64) if init_resid is not None else None
        if r is not None and r.shape != (self.op.n,):
            raise InvalidArgument(1, "init_resid has the wrong length")
        _check_op(self.op, lib().sb200_sym_init(self.h, _p(r)))

    def compute(self, selection=SortRule.LargestMagn, maxit: int = 1000, tol: float = 1e-10, sorting=SortRule.LargestAlge) -> int:
        nconv = C.c_int64()
        _check_op(self.op, lib().sb200_sym_compute(self.h, int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting), C.byref(nconv)))
        return nconv.value

    def info(self) -> CompInfo:
        v = C.c_int()
        _check(lib().sb200_sym_info(self.h, C.byref(v)))
        return CompInfo(v.value)

    def num_iterations(self) -> int:
        v = C.c_int64()
        _check(lib().sb200_sym_num_iterations(self.h, C.byref(v)))
        return v.value

    def num_operations(self) -> int:
        v = C.c_int64()
        _check(lib().sb200_sym_num_operations(self.h, C.byref(v)))
        return v.value

    def eigenvalues(self) -> np.ndarray:
        out = np.empty(self.nev)
        cnt = C.c_int64()
        _check(lib().sb200_sym_eigenvalues(self.h, _p(out), C.byref(cnt)))
        return out[:cnt.value].copy()

    def eigenvectors(self, nvec: int | None = None, local: bool = False, out: np.ndarray | None = None) -> np.ndarray:
        """n x nconv column-major matrix (HermEigsBase.h:447-470).  `out` may supply a (pinned) F-ordered buffer."""
        nvec = self.nev if nvec is None else int(nvec)
        rows = self.op.nrows_local if local else self.op.n
        if out is None:
            out = np.empty((rows, max(nvec, 1)), order="F")
        assert out.flags.f_contiguous and out.shape[0] == rows and out.shape[1] >= max(nvec, 1) and out.dtype == np.float64
        cnt = C.c_int64()
        fn = lib().sb200_sym_eigenvectors_local if local else lib().sb200_sym_eigenvectors
        _check(fn(self.h, C.c_int64(nvec), _p(out), C.byref(cnt)))
        return out[:, :cnt.value]

    def stats(self) -> dict:
        s = Stats()
        _check(lib().sb200_sym_stats(self.h, C.byref(s)))
        return s.as_dict()

    # ---- factorisation-tier hooks (test/Arnoldi.cpp) ----
    def factorize_from(self, from_k: int, to_m: int):
        _check(lib().sb200_sym_factorize_from(self.h, C.c_int64(from_k), C.c_int64(to_m)))

    def factorization(self):
        n, m = self.op.nrows_local, min(self.ncv, self.op.n)
        V = np.empty((n, m), order="F")
        H = np.empty((m, m), order="F")
        f = np.empty(n)
        beta, k = C.c_double(), C.c_int64()
        _check(lib().sb200_sym_get_factorization(self.h, _p(V), _p(H), _p(f), C.byref(beta), C.byref(k)))
        return dict(V=V, H=H, f=f, beta=beta.value, k=k.value)

    def close(self):
        if getattr(self, "h", None):
            lib().sb200_sym_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HermEigsSolver(SymEigsSolver):
    """HermEigsSolver.h:121-122 (HermEigsBase with a complex Scalar) on a SparseHermMatProd: real eigenvalues, complex vectors."""

    def _create(self, op, nev, ncv):
        _check(lib().sb200_herm_create(op.h, C.c_int64(nev), C.c_int64(ncv), C.byref(self.h)))

    def init(self, init_resid: np.ndarray | None = None):
        r = np.ascontiguousarray(init_resid, dtype=np.complex128) if init_resid is not None else None
        if r is not None and r.shape != (self.op.n,):
            raise InvalidArgument(1, "init_resid has the wrong length")
        _check_op(self.op, lib().sb200_sym_init(self.h, _p(r)))

    def eigenvectors(self, nvec: int | None = None) -> np.ndarray:
        nvec = self.nev if nvec is None else int(nvec)
        out = np.empty((self.op.n, max(nvec, 1)), dtype=np.complex128, order="F")
        cnt = C.c_int64()
        _check(lib().sb200_sym_eigenvectors(self.h, C.c_int64(nvec), _p(out), C.byref(cnt)))
        return out[:, :cnt.value]

    def factorization(self):
        n, m = self.op.n, min(self.ncv, self.op.n)
        V = np.empty((n, m), dtype=np.complex128, order="F")
        H = np.empty((m, m), order="F")
        f = np.empty(n, dtype=np.complex128)
        beta, k = C.c_double(), C.c_int64()
        _check(lib().sb200_sym_get_factorization(self.h, _p(V), _p(H), _p(f), C.byref(beta), C.byref(k)))
        return dict(V=V, H=H, f=f, beta=beta.value, k=k.value)


class SymEigsShiftSolver(SymEigsSolver):
    """SymEigsShiftSolver.h:148-196 — shift-and-invert mode: `op` computes (A - sigma I)^{-1} x (SparseSymShiftSolve, or a
    UserOp with a set_shift method); eigenvalues come back as lambda = 1/nu + sigma."""

    def __init__(self, op, nev: int, ncv: int, sigma: float):
        self.sigma = float(sigma)
        super().__init__(op, nev, ncv)
        if isinstance(op, UserOp):
            op.set_shift(self.sigma)

    def _create(self, op, nev, ncv):
        _check(lib().sb200_sym_create_shift(op.h, C.c_int64(nev), C.c_int64(ncv), C.c_double(self.sigma), C.byref(self.h)))


class GenEigsSolver:
    """GenEigsSolver.h:158-186 / GenEigsBase.h (real double matrices; complex Ritz pairs)."""

    def __init__(self, op: _SparseOp, nev: int, ncv: int):
        self.op = op
        self.nev, self.ncv = int(nev), int(ncv)
        self.h = C.c_void_p()
        _check(lib().sb200_gen_create(op.h, C.c_int64(nev), C.c_int64(ncv), C.byref(self.h)))

    def _complex_op(self) -> bool:
        """True for complex operators (SparseHermMatProd in 'general' mode, UserOp(complex_scalar=True)): only init() and the
        factorisation tier run for them in this build (the complex restart kernels are not built yet)."""
        return isinstance(self.op, SparseHermMatProd) or getattr(self.op, "_dtype", np.float64) == np.complex128

    def init(self, init_resid: np.ndarray | None = None):
        r = np.ascontiguousarray(init_resid, dtype=np.complex128 if self._complex_op() else np.float64) if init_resid is not None else None
        if r is not None and r.shape != (self.op.n,):
            raise InvalidArgument(1, "init_resid has the wrong length")
        _check_op(self.op, lib().sb200_gen_init(self.h, _p(r)))

    def compute(self, selection=SortRule.LargestMagn, maxit: int = 1000, tol: float = 1e-10, sorting=SortRule.LargestMagn) -> int:
        nconv = C.c_int64()
        _check_op(self.op, lib().sb200_gen_compute(self.h, int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting), C.byref(nconv)))
        return nconv.value

    def info(self) -> CompInfo:
        v = C.c_int()
        _check(lib().sb200_gen_info(self.h, C.byref(v)))
        return CompInfo(v.value)

    def num_iterations(self) -> int:
        v = C.c_int64()
        _check(lib().sb200_gen_num_iterations(self.h, C.byref(v)))
        return v.value

    def num_operations(self) -> int:
        v = C.c_int64()
        _check(lib().sb200_gen_num_operations(self.h, C.byref(v)))
        return v.value

    def eigenvalues(self) -> np.ndarray:
        out = np.empty(2 * self.nev)
        cnt = C.c_int64()
        _check(lib().sb200_gen_eigenvalues(self.h, _p(out), C.byref(cnt)))
        return (out[0::2] + 1j * out[1::2])[:cnt.value].copy()

    def eigenvectors(self, nvec: int | None = None) -> np.ndarray:
        nvec = self.nev if nvec is None else int(nvec)
        n = self.op.n
        out = np.empty(2 * n * max(nvec, 1))
        cnt = C.c_int64()
        _check(lib().sb200_gen_eigenvectors(self.h, C.c_int64(nvec), _p(out), C.byref(cnt)))
        Z = (out[0::2] + 1j * out[1::2]).reshape((n, max(nvec, 1)), order="F")
        return Z[:, :cnt.value]

    def stats(self) -> dict:
        s = Stats()
        _check(lib().sb200_gen_stats(self.h, C.byref(s)))
        return s.as_dict()

    def factorize_from(self, from_k: int, to_m: int):
        _check(lib().sb200_gen_factorize_from(self.h, C.c_int64(from_k), C.c_int64(to_m)))

    def factorization(self):
        n, m = self.op.nrows_local, min(self.ncv, self.op.n)
        dt = np.complex128 if self._complex_op() else np.float64
        V = np.empty((n, m), dtype=dt, order="F")
        H = np.empty((m, m), dtype=dt, order="F")
        f = np.empty(n, dtype=dt)
        beta, k = C.c_double(), C.c_int64()
        _check(lib().sb200_gen_get_factorization(self.h, _p(V), _p(H), _p(f), C.byref(beta), C.byref(k)))
        return dict(V=V, H=H, f=f, beta=beta.value, k=k.value)

    def close(self):
        if getattr(self, "h", None):
            lib().sb200_gen_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class dense:
    """Unit-tier hooks for the small dense device kernels (test/QR.cpp, test/Eigen.cpp)."""

    @staticmethod
    def _cm(H):
        return np.asfortranarray(np.asarray(H, dtype=np.float64))

    @staticmethod
    def tridiag_eigen(H):
        H = dense._cm(H)
        m = H.shape[0]
        ev = np.empty(m)
        Z = np.empty((m, m), order="F")
        _check(lib().sb200_dense_tridiag_eigen(C.c_int64(m), _p(H), _p(ev), _p(Z)))
        return ev, Z

    @staticmethod
    def givens(x, y, variant=0):
        """(r, c, s) arrays of Givens<double>::compute_rotation (Givens.h:166-205) evaluated on the device, one rotation per entry."""
        x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
        y = np.ascontiguousarray(np.atleast_1d(y), dtype=np.float64)
        if x.shape != y.shape:
            raise InvalidArgument(1, "x and y must have the same length")
        r, c, s = (np.empty(x.size) for _ in range(3))
        _check(lib().sb200_dense_givens(int(variant), C.c_int64(x.size), _p(x), _p(y), _p(r), _p(c), _p(s)))
        return r, c, s

    @staticmethod
    def shifted_qr(H, shift, kind="tridiag"):
        H = dense._cm(H)
        m = H.shape[0]
        D, Q = (np.empty((m, m), order="F") for _ in range(2))
        _check(lib().sb200_dense_shifted_qr(0 if kind == "tridiag" else 1, C.c_int64(m), _p(H), C.c_double(shift), _p(D), _p(Q)))
        return D, Q

    @staticmethod
    def double_shift_qr(H, s, t):
        H = dense._cm(H)
        m = H.shape[0]
        D, Q = (np.empty((m, m), order="F") for _ in range(2))
        _check(lib().sb200_dense_double_shift_qr(C.c_int64(m), _p(H), C.c_double(s), C.c_double(t), _p(D), _p(Q)))
        return D, Q

    @staticmethod
    def hess_eigen(H):
        H = dense._cm(H)
        m = H.shape[0]
        ev = np.empty(2 * m)
        V = np.empty(2 * m * m)
        _check(lib().sb200_dense_hess_eigen(C.c_int64(m), _p(H), _p(ev), _p(V)))
        return ev[0::2] + 1j * ev[1::2], (V[0::2] + 1j * V[1::2]).reshape((m, m), order="F")

    @staticmethod
    def shifted_qr_z(H, shift):
        """UpperHessenbergQR<complex>: (Q^H H Q, Q) for H - shift I = Q R, complex m x m (m <= 63)"""
        H = np.asfortranarray(H, dtype=np.complex128)
        m = H.shape[0]
        D, Q = np.empty((m, m), dtype=np.complex128, order="F"), np.empty((m, m), dtype=np.complex128, order="F")
        _check(lib().sb200_dense_shifted_qr_z(C.c_int64(m), _p(H), C.c_double(complex(shift).real), C.c_double(complex(shift).imag), _p(D), _p(Q)))
        return D, Q

    @staticmethod
    def hess_eigen_z(H):
        """UpperHessenbergEigen<complex>: eigenvalues (m) and unit-norm eigenvectors (m x m) of a complex Hessenberg matrix"""
        H = np.asfortranarray(H, dtype=np.complex128)
        m = H.shape[0]
        ev, Z = np.empty(m, dtype=np.complex128), np.empty((m, m), dtype=np.complex128, order="F")
        _check(lib().sb200_dense_hess_eigen_z(C.c_int64(m), _p(H), _p(ev), _p(Z)))
        return ev, Z

    @staticmethod
    def sym_restart(H, beta, nev, selection, tol):
        H = dense._cm(H)
        m = H.shape[0]
        rv, re = np.empty(m), np.empty(m)
        conv = np.empty(nev, np.int32)
        nconv, k = C.c_int64(), C.c_int64()
        Q, Hn = (np.empty((m, m), order="F") for _ in range(2))
        _check(lib().sb200_dense_sym_restart(C.c_int64(m), _p(H), C.c_double(beta), C.c_int64(nev), int(selection), C.c_double(tol), _p(rv), _p(re), _p(conv),
                                             C.byref(nconv), C.byref(k), _p(Q), _p(Hn)))
        return dict(ritz_val=rv, ritz_est=re, conv=conv, nconv=nconv.value, k=k.value, Q=Q, H=Hn)

    @staticmethod
    def compress(V, Q, kk, f=None, H=None, impl=0):
        """Restart GEMM (Arnoldi.h:320-340): returns V @ Q[:, :kk] and, when f is given, (Vnew, f_new, ||f_new||^2)."""
        V = dense._cm(V)
        Q = dense._cm(Q)
        n, m = V.shape
        out = np.empty((n, kk), order="F")
        if f is None:
            _check(lib().sb200_dense_compress(C.c_int64(n), C.c_int64(m), C.c_int64(kk), _p(V), _p(Q), None, _p(out), None, None, int(impl)))
            return out
        fb = np.array(f, dtype=np.float64, copy=True)
        Hc = dense._cm(H)
        nrm2 = C.c_double()
        _check(lib().sb200_dense_compress(C.c_int64(n), C.c_int64(m), C.c_int64(kk), _p(V), _p(Q), _p(Hc), _p(out), _p(fb), C.byref(nrm2), int(impl)))
        return out, fb, nrm2.value
