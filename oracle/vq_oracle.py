"""numpy/ctypes front end of oracle/vq_oracle.c (TEST INFRASTRUCTURE ONLY).

Mirrors the call structure of the reference quantizers
(/root/reference/models/archs/vqgan_arch.py:79-122, :212-287, :375-461).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvq_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "vq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libvq_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_SO)
        P, I, L = C.c_void_p, C.c_int, C.c_int64
        lib.vq_oracle_search.restype = I
        lib.vq_oracle_search.argtypes = [P, P, P, I, I, I, I, I, I, I, L, P, P, P, P, P]
        lib.vq_oracle_gather.restype = None
        lib.vq_oracle_gather.argtypes = [P, P, P, I, I, I, I, I, I, I, P]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def nearest_ids(mask, ht, wt):
    """F.interpolate(mask, (ht, wt), mode='nearest') on a float id map [B,1,Hs,Ws] -> int32 [B,ht,wt]
    (reference :222, :385-389); non-integer values select no codebook (the reference compares
    `segm_map == k` on floats, :243)."""
    mask = np.asarray(mask, dtype=np.float32)
    B, _, Hs, Ws = mask.shape
    sy = np.minimum(np.floor(np.arange(ht, dtype=np.float32) * np.float32(Hs / ht)).astype(np.int64), Hs - 1)
    sx = np.minimum(np.floor(np.arange(wt, dtype=np.float32) * np.float32(Ws / wt)).astype(np.int64), Ws - 1)
    v = mask[:, 0][:, sy][:, :, sx]
    ok = (v == np.floor(v)) & (v >= -1) & (v < 1e6)
    return np.where(ok, v, -1).astype(np.int32)


def search(z_nhwc, codebook, book_id, ps=1, cont_stride=None):
    """z_nhwc float32 [B,Hz,Wz,Cz]; codebook float32 [n_books,n_e,D]; book_id int32 [B,Hp,Wp] or None.
    Returns dict(idx, idx_cont, idx_list, zq_nhwc, sqerr)."""
    lib = _load()
    z = np.ascontiguousarray(z_nhwc, dtype=np.float32)
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    B, Hz, Wz, Cz = z.shape
    n_books, n_e, D = cb.shape
    assert D == Cz * ps * ps
    Hp, Wp = Hz // ps, Wz // ps
    bid = np.ascontiguousarray(book_id, dtype=np.int32) if book_id is not None else None
    idx = np.empty((B, Hp, Wp), np.int64)
    idx_cont = np.empty((B, Hp, Wp), np.int64)
    idx_list = np.empty((n_books, B, Hp, Wp), np.int64)
    zq = np.empty_like(z)
    err = C.c_double(0.0)
    rc = lib.vq_oracle_search(_p(z), _p(cb), _p(bid), B, Hz, Wz, Cz, ps, n_books, n_e,
                              n_e if cont_stride is None else cont_stride, _p(idx), _p(idx_cont),
                              _p(idx_list), _p(zq), C.byref(err))
    assert rc == 0
    return dict(idx=idx, idx_cont=idx_cont, idx_list=idx_list, zq_nhwc=zq, sqerr=err.value)


def gather(codebook, idx, book_id, B, Hz, Wz, Cz, ps=1):
    lib = _load()
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    n_books, n_e, _ = cb.shape
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    bid = np.ascontiguousarray(book_id, dtype=np.int32) if book_id is not None else None
    zq = np.empty((B, Hz, Wz, Cz), np.float32)
    lib.vq_oracle_gather(_p(cb), _p(idx), _p(bid), B, Hz, Wz, Cz, ps, n_books, n_e, _p(zq))
    return zq
