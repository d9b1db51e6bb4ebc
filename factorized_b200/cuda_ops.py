"""ctypes binding of libmfm_b200.so (include/mfm_b200.h) -- the only compute backend.

There is no CPU or eager-PyTorch fallback: if the shared library is missing, or a
tensor is not a CUDA fp32 tensor, this raises.  torch is used here for device
memory and the current stream only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmfm_b200.so")

GEMM_MODE = {"nt": 0, "nn": 1, "tn": 2}
PATH_SIMT_FP32, PATH_TC_BF16X3, PATH_TC_BF16 = 0, 1, 2

c_f = C.c_void_p   # device pointers travel as void*
LL = C.c_longlong


class LstmCell(C.Structure):
    _fields_ = [("T", C.c_int), ("B", C.c_int), ("h", C.c_int), ("gx_steps", C.c_int),
                ("gx", c_f), ("bias_rest", c_f), ("W", c_f),
                ("hs", c_f), ("ld_hs", LL), ("cs", c_f), ("ld_cs", LL), ("gates", c_f),
                ("dh_all", c_f), ("ld_dh_all", LL), ("dh_last", c_f), ("ld_dh_last", LL),
                ("dc_ext", c_f), ("ld_dc_ext", LL), ("dG", c_f), ("dc_scratch", c_f),
                ("cs_dup", c_f), ("dc_ext2", c_f), ("ld_gx", LL),
                ("dc_last", c_f), ("dh_out", c_f), ("dc_out", c_f), ("dc_ext2_full", C.c_int)]


class MemArgs(C.Structure):
    _fields_ = [("T", C.c_int), ("B", C.c_int), ("mem", C.c_int), ("g1", C.c_int), ("g2", C.c_int),
                ("G1pre", c_f), ("G2pre", c_f), ("cHat", c_f),
                ("W1m", c_f), ("ld_w1m", LL), ("W2m", c_f), ("ld_w2m", LL),
                ("W12", c_f), ("b12", c_f), ("W22", c_f), ("b22", c_f),
                ("mems", c_f), ("U1", c_f), ("U2", c_f), ("Gam1", c_f), ("Gam2", c_f),
                ("drop_p1", C.c_float), ("drop_p2", C.c_float), ("site1", C.c_int), ("site2", C.c_int), ("rng", c_f),
                ("scale1", C.c_float), ("scale2", C.c_float),
                ("dmem_last", c_f), ("ld_dmem_last", LL),
                ("dU1", c_f), ("dU2", c_f), ("dP1", c_f), ("dP2", c_f), ("dPc", c_f), ("ld_dU1", LL), ("ld_dU2", LL)]


_SIGS = {
    "mfm_version": (C.c_int, []),
    "mfm_launch_count": (C.c_ulonglong, []),
    "mfm_set_gemm_path": (C.c_int, [C.c_int]),
    "mfm_get_gemm_path": (C.c_int, []),
    "mfm_set_gemm_tc_min_work": (C.c_int, [LL]),
    "mfm_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_f, LL, c_f, LL, c_f, LL, c_f, c_f, C.c_int, C.c_int,
                           c_f, LL, C.c_float, C.c_float, C.c_int, c_f, c_f, c_f]),
    "mfm_gemm_ws": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_f, LL, c_f, LL, c_f, LL, c_f, c_f, C.c_int, C.c_int,
                              c_f, LL, C.c_float, C.c_float, C.c_int, c_f, c_f, c_f, LL, c_f]),
    "mfm_gemm_mse": (C.c_int, [C.c_int, C.c_int, C.c_int, c_f, LL, c_f, LL, c_f, c_f, LL, C.c_float, C.c_float, c_f, c_f, LL, c_f, LL,
                               c_f, LL, c_f]),
    "mfm_gemm_tn_pair": (C.c_int, [C.c_int, C.c_int, c_f, LL, C.c_int, c_f, LL, c_f, LL, c_f, C.c_int, c_f, LL, c_f, LL, c_f]),
    "mfm_debug_set_gemm_trace": (C.c_int, [c_f, LL]),
    "mfm_debug_stamp": (C.c_int, [c_f, C.c_int, c_f]),
    "mfm_lstm_seq_fwd": (C.c_int, [C.POINTER(LstmCell), C.c_int, c_f]),
    "mfm_lstm_seq_bwd": (C.c_int, [C.POINTER(LstmCell), C.c_int, c_f]),
    "mfm_debug_lstm_force_nb": (C.c_int, [C.c_int]),
    "mfm_debug_lstm_force_chains": (C.c_int, [C.c_int]),
    "mfm_debug_set_lstm_trace": (C.c_int, [c_f]),
    "mfm_debug_lstm_variant_count": (C.c_ulonglong, [C.c_int]),
    "mfm_debug_gemm_ps_count": (C.c_int, []),
    "mfm_debug_gemm_ps_residency": (C.c_int, [C.c_int]),
    "mfm_debug_set_gemm_ps_trace": (C.c_int, [c_f]),
    "mfm_debug_mem_force_simt": (C.c_int, [C.c_int]),
    "mfm_debug_mem_ws_trace": (C.c_int, [c_f]),
    "mfm_debug_mem_ws_count": (C.c_ulonglong, [C.c_int]),
    "mfm_mfn_mem_fwd": (C.c_int, [C.POINTER(MemArgs), c_f]),
    "mfm_mfn_mem_bwd": (C.c_int, [C.POINTER(MemArgs), c_f]),
    "mfm_softmax_gate_fwd": (C.c_int, [C.c_int, C.c_int, c_f, c_f, c_f, c_f]),
    "mfm_softmax_gate_bwd": (C.c_int, [C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, c_f]),
    "mfm_mmd_fwd": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, LL, c_f, c_f]),
    "mfm_mmd_bwd": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, LL, C.c_float, c_f, c_f, LL, c_f]),
    "mfm_randn": (C.c_int, [LL, c_f, c_f, C.c_int, c_f]),
    "mfm_rownorm2": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, c_f]),
    "mfm_mmd_kexp": (C.c_int, [C.c_int, C.c_int, c_f, c_f, c_f, C.c_int, C.c_float, c_f, c_f]),
    "mfm_mmd_kexp64": (C.c_int, [C.c_int, C.c_int, c_f, c_f, c_f, C.c_int, C.c_double, c_f, c_f]),
    "mfm_mmd_fold": (C.c_int, [C.c_int, c_f, c_f, c_f]),
    "mfm_mmd_combine": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, c_f, c_f, c_f, C.c_float, c_f, c_f, LL, c_f]),
    "mfm_copy2d": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, LL, C.c_int, c_f]),
    "mfm_add": (C.c_int, [LL, c_f, c_f, c_f, c_f]),
    "mfm_zero": (C.c_int, [LL, c_f, c_f]),
    "mfm_colsum": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, c_f]),
    "mfm_relu_bwd": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, LL, c_f, LL, c_f]),
    "mfm_mse_fwd_bwd": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, LL, C.c_float, C.c_float, c_f, c_f, LL, c_f]),
    "mfm_l1_fwd_bwd": (C.c_int, [LL, c_f, c_f, C.c_float, c_f, c_f, c_f]),
    "mfm_ce_fwd_bwd": (C.c_int, [C.c_int, C.c_int, c_f, c_f, C.c_float, c_f, c_f, c_f]),
    "mfm_kld_fwd": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, LL, c_f, c_f]),
    "mfm_kld_bwd": (C.c_int, [C.c_int, C.c_int, c_f, LL, c_f, LL, C.c_float, c_f, c_f, LL, c_f, LL, c_f]),
    "mfm_loss_total": (C.c_int, [c_f, C.c_float, C.c_float, C.c_float, C.c_float, c_f]),
    "mfm_adam_step": (C.c_int, [LL, c_f, c_f, c_f, c_f, c_f, C.c_float, C.c_double, C.c_double, C.c_double, c_f]),
    "mfm_rng_tick": (C.c_int, [c_f, c_f]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libmfm_b200.so and declare every prototype of include/mfm_b200.h.  Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "factorized_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or eager fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class MfmCudaError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        if rc < 0:
            raise MfmCudaError("%s: bad argument / unsupported (code %d)" % (what, rc))
        raise MfmCudaError("%s: CUDA error %d (%s)" % (what, rc, _cuda_err_name(rc)))


def _cuda_err_name(rc: int) -> str:
    try:
        rt = C.CDLL("libcudart.so")
        rt.cudaGetErrorString.restype = C.c_char_p
        return rt.cudaGetErrorString(rc).decode()
    except Exception:
        return "?"


def _mat(t: torch.Tensor, what: str):
    """(ptr, rows, cols, ld) of a 2-D fp32 CUDA tensor with unit column stride."""
    if not t.is_cuda:
        raise MfmCudaError("%s: tensor is on %s; factorized_b200 runs on CUDA only (no CPU fallback)" % (what, t.device))
    if t.dtype != torch.float32:
        raise MfmCudaError("%s: dtype %s, need float32" % (what, t.dtype))
    if t.dim() != 2:
        raise MfmCudaError("%s: need a 2-D matrix, got %s" % (what, tuple(t.shape)))
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise MfmCudaError("%s: column stride %d, need 1" % (what, t.stride(1)))
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return t.data_ptr(), t.shape[0], t.shape[1], ld


def _vec(t: Optional[torch.Tensor], what: str, n: Optional[int] = None):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32:
        raise MfmCudaError("%s: need a CUDA float32 vector" % what)
    if not t.is_contiguous():
        raise MfmCudaError("%s: vector must be contiguous" % what)
    if n is not None and t.numel() != n:
        raise MfmCudaError("%s: expected %d elements, got %d" % (what, n, t.numel()))
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CudaOps:
    """The primitive set the Engine schedules, bound to the sm_100a kernels."""

    def __init__(self):
        self.lib = load_library()
        self._ws = {}
        if not torch.cuda.is_available():
            raise RuntimeError("factorized_b200 needs a CUDA device (B200, sm_100a); none is visible")

    @property
    def launches(self) -> int:
        return int(self.lib.mfm_launch_count())

    def set_gemm_path(self, path: int, min_work: Optional[int] = None):
        _check(self.lib.mfm_set_gemm_path(path), "mfm_set_gemm_path")
        if min_work is not None:
            _check(self.lib.mfm_set_gemm_tc_min_work(int(min_work)), "mfm_set_gemm_tc_min_work")

    def get_gemm_path(self) -> int:
        return int(self.lib.mfm_get_gemm_path())

    # ---- GEMM ----
    def gemm(self, mode, A, B, Cm, bias=None, bias2=None, act=0, accumulate=False, mask=None, mask_scale=1.0,
             drop=None, rng=None, colsum_out=None):
        pa, ar, ac, lda = _mat(A, "gemm A")
        pb, br, bc, ldb = _mat(B, "gemm B")
        pc, M, N, ldc = _mat(Cm, "gemm C")
        if mode == "nt":
            K = ac
            ok = (ar == M and br == N and bc == K)
        elif mode == "nn":
            K = ac
            ok = (ar == M and br == K and bc == N)
        elif mode == "tn":
            K = ar
            ok = (ac == M and br == K and bc == N)
        else:
            raise ValueError(mode)
        if not ok:
            raise MfmCudaError("gemm %s: shapes A%s B%s C%s do not agree" % (mode, tuple(A.shape), tuple(B.shape), tuple(Cm.shape)))
        pm, ldm = None, 0
        if mask is not None:
            pm, mr, mc, ldm = _mat(mask, "gemm mask")
            if (mr, mc) != (M, N):
                raise MfmCudaError("gemm mask shape")
        p, site = (0.0, 0) if drop is None else drop
        prng = None
        if drop is not None:
            if rng is None or rng.dtype != torch.int64 or not rng.is_cuda:
                raise MfmCudaError("dropout needs a CUDA int64 rng state [seed, step]")
            prng = rng.data_ptr()
        ws, ws_bytes = None, 0
        if mode != "tn" and M >= 4096:          # weight operand against many row tiles: pre-split it once per call
            ws_bytes = 8 * (N + 64) * (K + 16)
            ws = self._gemm_workspace(Cm.device, ws_bytes)
        _check(self.lib.mfm_gemm_ws(GEMM_MODE[mode], M, N, K, pa, lda, pb, ldb, pc, ldc, _vec(bias, "bias", N),
                                    _vec(bias2, "bias2", N), act, int(accumulate), pm, ldm, float(mask_scale),
                                    float(p), int(site), prng,
                                    None if colsum_out is None else _vec(colsum_out, "colsum_out", M),
                                    ws, ws_bytes, _stream()), "mfm_gemm_ws")

    def gemm_mse(self, A, W, bias, x, loss_scale, grad_scale, slot, dxhat, xhat=None):
        """x_hat = A W^T + bias fused with the MSE head: slot += loss_scale * sum (x_hat - x)^2, dxhat = grad_scale * (x_hat - x);
        x_hat is written only when `xhat` is given."""
        pa, M, K, lda = _mat(A, "gemm_mse A")
        pw, N, K2, ldw = _mat(W, "gemm_mse W")
        px, mx, nx, ldx = _mat(x, "gemm_mse x")
        pd, md, nd, ldd = _mat(dxhat, "gemm_mse dxhat")
        if K2 != K or (mx, nx) != (M, N) or (md, nd) != (M, N):
            raise MfmCudaError("gemm_mse: shapes A%s W%s x%s dxhat%s do not agree" % (tuple(A.shape), tuple(W.shape), tuple(x.shape), tuple(dxhat.shape)))
        ph, ldh = None, 0
        if xhat is not None:
            ph, mh, nh, ldh = _mat(xhat, "gemm_mse xhat")
            if (mh, nh) != (M, N):
                raise MfmCudaError("gemm_mse xhat shape")
        ws, ws_bytes = None, 0
        if M >= 4096:
            ws_bytes = 8 * (N + 64) * (K + 16)
            ws = self._gemm_workspace(dxhat.device, ws_bytes)
        _check(self.lib.mfm_gemm_mse(M, N, K, pa, lda, pw, ldw, _vec(bias, "bias", N), px, ldx, float(loss_scale), float(grad_scale),
                                     _vec(slot, "mse slot", 1), pd, ldd, ph, ldh, ws, ws_bytes, _stream()), "mfm_gemm_mse")

    def gemm_tn_pair(self, dY, A1, C1, colsum1, A2, C2):
        """C1 += dY^T A1 (colsum1 += column sums of dY, may be None) and C2 += dY^T A2 in one launch."""
        pa, K, M, lda = _mat(dY, "gemm_tn_pair dY")
        p1, k1, N1, ld1 = _mat(A1, "gemm_tn_pair A1")
        p2, k2, N2, ld2 = _mat(A2, "gemm_tn_pair A2")
        c1, m1, n1, lc1 = _mat(C1, "gemm_tn_pair C1")
        c2, m2, n2, lc2 = _mat(C2, "gemm_tn_pair C2")
        if (k1, k2, m1, m2, n1, n2) != (K, K, M, M, N1, N2):
            raise MfmCudaError("gemm_tn_pair: shapes do not agree")
        _check(self.lib.mfm_gemm_tn_pair(M, K, pa, lda, N1, p1, ld1, c1, lc1,
                                         None if colsum1 is None else _vec(colsum1, "colsum1", M),
                                         N2, p2, ld2, c2, lc2, _stream()), "mfm_gemm_tn_pair")

    def stamp(self, buf: torch.Tensor, slot: int):
        """Debug: write the device clock to buf[slot] (int64) when the current stream gets here."""
        _check(self.lib.mfm_debug_stamp(buf.data_ptr(), int(slot), _stream()), "mfm_debug_stamp")

    def _gemm_workspace(self, device, nbytes: int) -> int:
        """One scratch buffer per (device, stream): calls on a stream are ordered, so the next call may reuse it."""
        key = (device.index, _stream())
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(nbytes, 4 << 20), dtype=torch.uint8, device=device)
            self._ws[key] = t
        return t.data_ptr()

    # ---- LSTM ----
    @staticmethod
    def _cells(cells, bwd):
        arr = (LstmCell * len(cells))()
        for i, c in enumerate(cells):
            s = arr[i]
            s.T, s.B, s.h = c["T"], c["B"], c["h"]
            h4 = 4 * c["h"]
            pW, wr, wc, ldw = _mat(c["W"], "lstm W")
            if (wr, wc) != (h4, c["h"]) or (wr > 1 and ldw != c["h"]):
                raise MfmCudaError("lstm W must be contiguous [4h,h]")
            s.W = pW
            pg, gr, gc, ldg = _mat(c["gates"], "lstm gates")
            if (gr, gc) != (c["T"] * c["B"], h4) or ldg != h4:
                raise MfmCudaError("lstm gates must be contiguous [T*B,4h]")
            s.gates = pg
            pcs, cr, cc, ldcs = _mat(c["cs"], "lstm cs")
            if (cr, cc) != ((c["T"] + 1) * c["B"], c["h"]):
                raise MfmCudaError("lstm cs must be [(T+1)*B,h]")
            s.cs, s.ld_cs = pcs, ldcs
            if not bwd:
                s.gx_steps = c["gx_steps"]
                pgx, xr, xc, ldx = _mat(c["gx"], "lstm gx")
                if (xr, xc) != (c["gx_steps"] * c["B"], h4):
                    raise MfmCudaError("lstm gx must be [gx_steps*B,4h]")
                s.gx, s.ld_gx = pgx, ldx
                s.bias_rest = _vec(c.get("bias_rest"), "lstm bias_rest", h4)
                if c["gx_steps"] < c["T"] and s.bias_rest is None:
                    raise MfmCudaError("lstm: bias_rest required when gx_steps < T")
                phs, hr, hc, ldhs = _mat(c["hs"], "lstm hs")
                if (hr, hc) != ((c["T"] + 1) * c["B"], c["h"]):
                    raise MfmCudaError("lstm hs must be [(T+1)*B,h]")
                s.hs, s.ld_hs = phs, ldhs
                if c.get("cs_dup") is not None:
                    p_, r_, c_, l_ = _mat(c["cs_dup"], "lstm cs_dup")
                    if (r_, c_) != ((c["T"] + 1) * c["B"], c["h"]) or l_ != ldcs:
                        raise MfmCudaError("lstm cs_dup must be [(T+1)*B,h] with the leading dimension of cs")
                    s.cs_dup = p_
            else:
                pd, dr, dc_, ldd = _mat(c["dG"], "lstm dG")
                if (dr, dc_) != (c["T"] * c["B"], h4) or ldd != h4:
                    raise MfmCudaError("lstm dG must be contiguous [T*B,4h]")
                s.dG = pd
                if c.get("dc_scratch") is not None:
                    p_, r_, c_, l_ = _mat(c["dc_scratch"], "lstm dc_scratch")
                    if (r_, c_) != (c["B"], c["h"]) or (r_ > 1 and l_ != c["h"]):
                        raise MfmCudaError("lstm dc_scratch must be contiguous [B,h]")
                    s.dc_scratch = p_
                if c.get("dh_all") is not None:
                    p_, r_, c_, l_ = _mat(c["dh_all"], "lstm dh_all")
                    if (r_, c_) != (c["T"] * c["B"], c["h"]):
                        raise MfmCudaError("lstm dh_all shape")
                    s.dh_all, s.ld_dh_all = p_, l_
                if c.get("dh_last") is not None:
                    p_, r_, c_, l_ = _mat(c["dh_last"], "lstm dh_last")
                    if (r_, c_) != (c["B"], c["h"]):
                        raise MfmCudaError("lstm dh_last shape")
                    s.dh_last, s.ld_dh_last = p_, l_
                if c.get("dc_ext") is not None:
                    p_, r_, c_, l_ = _mat(c["dc_ext"], "lstm dc_ext")
                    if (r_, c_) != (c["T"] * c["B"], c["h"]):
                        raise MfmCudaError("lstm dc_ext shape")
                    s.dc_ext, s.ld_dc_ext = p_, l_
                full2 = bool(c.get("dc_ext2_full"))
                if c.get("dc_ext2") is not None and (c["T"] > 1 or full2):
                    p_, r_, c_, l_ = _mat(c["dc_ext2"], "lstm dc_ext2")
                    want = (c["T"] if full2 else c["T"] - 1) * c["B"]
                    if c.get("dc_ext") is None or (r_, c_) != (want, c["h"]) or l_ != s.ld_dc_ext:
                        raise MfmCudaError("lstm dc_ext2 must be [(T-1)*B,h] ([T*B,h] with dc_ext2_full) with the leading dimension of dc_ext")
                    s.dc_ext2 = p_
                    s.dc_ext2_full = 1 if full2 else 0
                for key in ("dc_last", "dh_out", "dc_out"):           # time-split recurrence: carried state in / out
                    if c.get(key) is not None:
                        p_, r_, c_, l_ = _mat(c[key], "lstm " + key)
                        if (r_, c_) != (c["B"], c["h"]) or (r_ > 1 and l_ != c["h"]):
                            raise MfmCudaError("lstm %s must be contiguous [B,h]" % key)
                        setattr(s, key, p_)
                if (c.get("dh_out") is None) != (c.get("dc_out") is None):
                    raise MfmCudaError("lstm dh_out and dc_out go together")
        return arr

    def lstm_fwd(self, cells):
        arr = self._cells(cells, False)
        _check(self.lib.mfm_lstm_seq_fwd(arr, len(cells), _stream()), "mfm_lstm_seq_fwd")

    def lstm_bwd(self, cells):
        arr = self._cells(cells, True)
        _check(self.lib.mfm_lstm_seq_bwd(arr, len(cells), _stream()), "mfm_lstm_seq_bwd")

    # ---- MFN memory ----
    @staticmethod
    def _mem(a, bwd):
        s = MemArgs()
        s.T, s.B, s.mem, s.g1, s.g2 = a["T"], a["B"], a["mem"], a["g1"], a["g2"]
        TB = a["T"] * a["B"]

        def dense(key, rows, cols):
            p, r, c, ld = _mat(a[key], "mem " + key)
            if (r, c) != (rows, cols) or (r > 1 and ld != cols):
                raise MfmCudaError("mem %s must be contiguous [%d,%d], got %s ld %d" % (key, rows, cols, tuple(a[key].shape), ld))
            return p
        s.cHat = dense("cHat", TB, a["mem"])
        s.W1m, _, _, s.ld_w1m = _mat(a["W1m"], "mem W1m")
        s.W2m, _, _, s.ld_w2m = _mat(a["W2m"], "mem W2m")
        if tuple(a["W1m"].shape) != (a["g1"], a["mem"]) or tuple(a["W2m"].shape) != (a["g2"], a["mem"]):
            raise MfmCudaError("mem W*m shapes")
        s.W12 = dense("W12", a["mem"], a["g1"])
        s.W22 = dense("W22", a["mem"], a["g2"])
        s.mems = dense("mems", (a["T"] + 1) * a["B"], a["mem"])
        s.U1, s.U2 = dense("U1", TB, a["g1"]), dense("U2", TB, a["g2"])
        s.Gam1, s.Gam2 = dense("Gam1", TB, a["mem"]), dense("Gam2", TB, a["mem"])
        if not bwd:
            s.G1pre, s.G2pre = dense("G1pre", TB, a["g1"]), dense("G2pre", TB, a["g2"])
            s.b12, s.b22 = _vec(a["b12"], "b12", a["mem"]), _vec(a["b22"], "b22", a["mem"])
            for k, (pk, sk) in (("drop1", ("drop_p1", "site1")), ("drop2", ("drop_p2", "site2"))):
                if a.get(k) is not None:
                    setattr(s, pk, float(a[k][0]))
                    setattr(s, sk, int(a[k][1]))
            if a.get("drop1") is not None or a.get("drop2") is not None:
                s.rng = a["rng"].data_ptr()
        else:
            s.scale1, s.scale2 = float(a["scale1"]), float(a["scale2"])
            s.dmem_last, _, _, s.ld_dmem_last = _mat(a["dmem_last"], "mem dmem_last")
            s.dU1, r1, c1, s.ld_dU1 = _mat(a["dU1"], "mem dU1")
            s.dU2, r2, c2, s.ld_dU2 = _mat(a["dU2"], "mem dU2")
            if (r1, c1) != (TB, a["g1"]) or (r2, c2) != (TB, a["g2"]):
                raise MfmCudaError("mem dU1/dU2 must be [T*B, g]")
            s.dP1, s.dP2, s.dPc = dense("dP1", TB, a["mem"]), dense("dP2", TB, a["mem"]), dense("dPc", TB, a["mem"])
        return s

    def mfn_mem_fwd(self, a):
        s = self._mem(a, False)
        _check(self.lib.mfm_mfn_mem_fwd(C.byref(s), _stream()), "mfm_mfn_mem_fwd")

    def mfn_mem_bwd(self, a):
        s = self._mem(a, True)
        _check(self.lib.mfm_mfn_mem_bwd(C.byref(s), _stream()), "mfm_mfn_mem_bwd")

    # ---- attention gate ----
    def softmax_gate_fwd(self, L, cstar, attended):
        p, M, N, ld = _mat(L, "softmax L")
        pc, _, _, ldc = _mat(cstar, "softmax cstar")
        pa, _, _, lda = _mat(attended, "softmax attended")
        if not (ld == ldc == lda == N) or cstar.shape != L.shape or attended.shape != L.shape:
            raise MfmCudaError("softmax_gate_fwd: contiguous equal shapes required")
        _check(self.lib.mfm_softmax_gate_fwd(M, N, p, pc, pa, _stream()), "mfm_softmax_gate_fwd")

    def softmax_gate_bwd(self, dAttended, att, cstar, dL, dcstar):
        ps = []
        M, N = att.shape
        for t, nm in ((dAttended, "dAttended"), (att, "att"), (cstar, "cstar"), (dL, "dL"), (dcstar, "dcstar")):
            p, r, c, ld = _mat(t, "softmax " + nm)
            if (r, c) != (M, N) or ld != N:
                raise MfmCudaError("softmax_gate_bwd: contiguous equal shapes required (%s)" % nm)
            ps.append(p)
        _check(self.lib.mfm_softmax_gate_bwd(M, N, *ps, _stream()), "mfm_softmax_gate_bwd")

    # ---- MMD ----
    def mmd_fwd(self, z, g, out):
        pz, B, dim, ldz = _mat(z, "mmd z")
        pg, gb, gd, ldg = _mat(g, "mmd g")
        if (gb, gd) != (B, dim):
            raise MfmCudaError("mmd: noise shape %s != latent shape %s" % (tuple(g.shape), tuple(z.shape)))
        _check(self.lib.mfm_mmd_fwd(B, dim, pz, ldz, pg, ldg, _vec(out, "mmd out", 1), _stream()), "mfm_mmd_fwd")

    def mmd_bwd(self, z, g, scale, dz, scale_dev=None):
        pz, B, dim, ldz = _mat(z, "mmd z")
        pg, gb, gd, ldg = _mat(g, "mmd g")
        pd, db, dd, ldd = _mat(dz, "mmd dz")
        if (gb, gd) != (B, dim) or (db, dd) != (B, dim):
            raise MfmCudaError("mmd_bwd shapes")
        psd = None
        if scale_dev is not None:
            if not scale_dev.is_cuda or scale_dev.dtype != torch.float32 or scale_dev.numel() != 1:
                raise MfmCudaError("mmd_bwd: scale_dev must be a CUDA float32 scalar")
            psd = scale_dev.data_ptr()
        _check(self.lib.mfm_mmd_bwd(B, dim, pz, ldz, pg, ldg, float(scale), psd, pd, ldd, _stream()), "mfm_mmd_bwd")

    def rownorm2(self, x, out):
        px, B, dim, ld = _mat(x, "rownorm2 x")
        _check(self.lib.mfm_rownorm2(B, dim, px, ld, _vec(out, "rownorm2 out", B), _stream()), "mfm_rownorm2")

    def mmd_kexp(self, S, nx, ny, dim, weight, slot):
        ps, M, N, ld = _mat(S, "mmd_kexp S")
        if ld != N:
            raise MfmCudaError("mmd_kexp: S must be contiguous")
        _check(self.lib.mfm_mmd_kexp(M, N, ps, _vec(nx, "nx", M), _vec(ny, "ny", N), int(dim), float(weight),
                                     _vec(slot, "slot", 1), _stream()), "mfm_mmd_kexp")

    def mmd_kexp64(self, S, nx, ny, dim, weight, acc):
        """mmd_kexp with a float64 accumulator element (acc: 1-element float64 CUDA tensor view)."""
        ps, M, N, ld = _mat(S, "mmd_kexp S")
        if ld != N:
            raise MfmCudaError("mmd_kexp: S must be contiguous")
        if not acc.is_cuda or acc.dtype != torch.float64 or acc.numel() != 1:
            raise MfmCudaError("mmd_kexp64: acc must be one CUDA float64 element")
        _check(self.lib.mfm_mmd_kexp64(M, N, ps, _vec(nx, "nx", M), _vec(ny, "ny", N), int(dim), float(weight),
                                       acc.data_ptr(), _stream()), "mfm_mmd_kexp64")

    def mmd_fold(self, acc, slots):
        if not acc.is_cuda or acc.dtype != torch.float64 or not acc.is_contiguous() or acc.numel() != slots.numel():
            raise MfmCudaError("mmd_fold: acc must be a contiguous CUDA float64 vector as long as slots")
        _check(self.lib.mfm_mmd_fold(acc.numel(), acc.data_ptr(), _vec(slots, "slots", acc.numel()), _stream()), "mfm_mmd_fold")

    def mmd_combine(self, z, rs, cs, t1, t2, scale, dz, scale_dev=None):
        pz, B, dim, ldz = _mat(z, "mmd_combine z")
        pd, B2, d2, ldd = _mat(dz, "mmd_combine dz")
        p1, _, _, l1 = _mat(t1, "mmd_combine t1")
        p2, _, _, l2 = _mat(t2, "mmd_combine t2")
        if (B2, d2) != (B, dim) or tuple(t1.shape) != (B, dim) or tuple(t2.shape) != (B, dim) or l1 != dim or l2 != dim:
            raise MfmCudaError("mmd_combine shapes")
        psd = None if scale_dev is None else scale_dev.data_ptr()
        _check(self.lib.mfm_mmd_combine(B, dim, pz, ldz, _vec(rs, "rs", B), _vec(cs, "cs", B), p1, p2, float(scale), psd,
                                        pd, ldd, _stream()), "mfm_mmd_combine")

    # ---- small kernels ----
    def copy2d(self, src, dst, accumulate=False):
        ps, M, N, lds = _mat(src, "copy2d src")
        pd, M2, N2, ldd = _mat(dst, "copy2d dst")
        if (M, N) != (M2, N2):
            raise MfmCudaError("copy2d shapes %s vs %s" % (tuple(src.shape), tuple(dst.shape)))
        _check(self.lib.mfm_copy2d(M, N, ps, lds, pd, ldd, int(accumulate), _stream()), "mfm_copy2d")

    def add(self, a, b, out):
        if not (a.is_contiguous() and b.is_contiguous() and out.is_contiguous()) or a.numel() != out.numel() or b.numel() != out.numel():
            raise MfmCudaError("add: contiguous equal-sized tensors required")
        for t in (a, b, out):
            if not t.is_cuda or t.dtype != torch.float32:
                raise MfmCudaError("add: CUDA float32 required")
        _check(self.lib.mfm_add(out.numel(), a.data_ptr(), b.data_ptr(), out.data_ptr(), _stream()), "mfm_add")

    def zero(self, t):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise MfmCudaError("zero: contiguous CUDA float32 required")
        _check(self.lib.mfm_zero(t.numel(), t.data_ptr(), _stream()), "mfm_zero")

    def colsum(self, A, out):
        pa, M, N, lda = _mat(A, "colsum A")
        _check(self.lib.mfm_colsum(M, N, pa, lda, _vec(out, "colsum out", N), _stream()), "mfm_colsum")

    def relu_bwd(self, dy, y, out):
        p1, M, N, l1 = _mat(dy, "relu_bwd dy")
        p2, M2, N2, l2 = _mat(y, "relu_bwd y")
        p3, M3, N3, l3 = _mat(out, "relu_bwd out")
        if not ((M, N) == (M2, N2) == (M3, N3)):
            raise MfmCudaError("relu_bwd shapes")
        _check(self.lib.mfm_relu_bwd(M, N, p1, l1, p2, l2, p3, l3, _stream()), "mfm_relu_bwd")

    def mse_fwd_bwd(self, xhat, x, loss_scale, grad_scale, slot, dxhat):
        p1, M, N, l1 = _mat(xhat, "mse xhat")
        p2, M2, N2, l2 = _mat(x, "mse x")
        if (M, N) != (M2, N2):
            raise MfmCudaError("mse shapes")
        p3, l3 = None, 0
        if dxhat is not None:
            p3, M3, N3, l3 = _mat(dxhat, "mse dxhat")
            if (M3, N3) != (M, N):
                raise MfmCudaError("mse dxhat shape")
        _check(self.lib.mfm_mse_fwd_bwd(M, N, p1, l1, p2, l2, float(loss_scale), float(grad_scale),
                                        _vec(slot, "mse slot", 1), p3, l3, _stream()), "mfm_mse_fwd_bwd")

    def l1_fwd_bwd(self, yhat, y, scale, slot, dy):
        if not (yhat.is_contiguous() and y.is_contiguous() and dy.is_contiguous()) or y.numel() != yhat.numel():
            raise MfmCudaError("l1: contiguous equal-sized tensors required")
        if y.dtype != torch.float32 or not y.is_cuda:
            raise MfmCudaError("l1: targets must be CUDA float32")
        _check(self.lib.mfm_l1_fwd_bwd(yhat.numel(), yhat.data_ptr(), y.data_ptr(), float(scale),
                                       _vec(slot, "l1 slot", 1), dy.data_ptr(), _stream()), "mfm_l1_fwd_bwd")

    def ce_fwd_bwd(self, yhat, y, scale, slot, dy):
        p, B, Cn, ld = _mat(yhat, "ce yhat")
        if ld != Cn or not dy.is_contiguous() or y.dtype != torch.int64 or not y.is_cuda or y.numel() != B:
            raise MfmCudaError("ce: contiguous logits and CUDA int64 labels [B] required")
        _check(self.lib.mfm_ce_fwd_bwd(B, Cn, p, y.data_ptr(), float(scale), _vec(slot, "ce slot", 1), dy.data_ptr(),
                                       _stream()), "mfm_ce_fwd_bwd")

    def kld_fwd(self, mu, logvar, slot):
        pm, M, N, ldm = _mat(mu, "kld mu")
        pl, M2, N2, ldl = _mat(logvar, "kld logvar")
        if (M, N) != (M2, N2):
            raise MfmCudaError("kld shapes")
        _check(self.lib.mfm_kld_fwd(M, N, pm, ldm, pl, ldl, _vec(slot, "kld slot", 1), _stream()), "mfm_kld_fwd")

    def kld_bwd(self, mu, logvar, scale, dmu, dlogvar, scale_dev=None):
        pm, M, N, ldm = _mat(mu, "kld mu")
        pl, M2, N2, ldl = _mat(logvar, "kld logvar")
        pdm, M3, N3, lddm = _mat(dmu, "kld dmu")
        pdl, M4, N4, lddl = _mat(dlogvar, "kld dlogvar")
        if not ((M, N) == (M2, N2) == (M3, N3) == (M4, N4)):
            raise MfmCudaError("kld shapes")
        psd = None if scale_dev is None else scale_dev.data_ptr()
        _check(self.lib.mfm_kld_bwd(M, N, pm, ldm, pl, ldl, float(scale), psd, pdm, lddm, pdl, lddl, _stream()), "mfm_kld_bwd")

    def loss_total(self, lb, l0, l1, l2, lmmd):
        _check(self.lib.mfm_loss_total(lb.data_ptr(), float(l0), float(l1), float(l2), float(lmmd), _stream()),
               "mfm_loss_total")

    def adam(self, p, g, m, v, state, grad_scale=1.0, betas=(0.9, 0.999), eps=1e-8):
        n = p.numel()
        for t in (p, g, m, v):
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
                raise MfmCudaError("adam: four contiguous CUDA float32 buffers of equal size required")
        _check(self.lib.mfm_adam_step(n, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), state.data_ptr(),
                                      float(grad_scale), float(betas[0]), float(betas[1]), float(eps), _stream()),
               "mfm_adam_step")

    def randn(self, out, rng, site):
        if not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous():
            raise MfmCudaError("randn: contiguous CUDA float32 required")
        _check(self.lib.mfm_randn(out.numel(), out.data_ptr(), rng.data_ptr(), int(site), _stream()), "mfm_randn")

    def rng_tick(self, rng):
        _check(self.lib.mfm_rng_tick(rng.data_ptr(), _stream()), "mfm_rng_tick")
