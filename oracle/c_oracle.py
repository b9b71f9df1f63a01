"""ctypes front-end to oracle/libhansel_oracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libhansel_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libhansel_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        p = C.c_void_p
        L.or_ingest.argtypes = [p, p, p, C.c_int64, C.c_int32, C.c_int32, p, p]
        L.or_ingest.restype = C.c_int
        L.or_u32_to_f32.argtypes = [p, p, C.c_int64]
        L.or_counts_all.argtypes = [p, C.c_int32, C.c_int32, p]
        L.or_edge_weights_at.argtypes = [p, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int,
                                         C.c_int32, p, p, p]
        L.or_edge_weights_at.restype = C.c_int
        L.or_generate_path.argtypes = [p, p, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int, p, p]
        L.or_generate_path.restype = C.c_int
        L.or_reweight_path.argtypes = [p, C.c_int32, C.c_int32, p, C.c_double]
        L.or_reweight_path.restype = C.c_double
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def ingest(rank, off, codes, n_snps, W):
    """-> (band uint32 [N+2][W][7][7], totals int64[4] = slices, crumbs, covered, sentinels)."""
    rank = np.ascontiguousarray(rank, dtype=np.int32)
    off = np.ascontiguousarray(off, dtype=np.int64)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    band = np.zeros((n_snps + 2, W, 7, 7), dtype=np.uint32)
    totals = np.zeros(4, dtype=np.int64)
    rc = lib().or_ingest(_ptr(rank), _ptr(off), _ptr(codes), len(rank), n_snps, W, _ptr(band), _ptr(totals))
    if rc != 0:
        raise ValueError("or_ingest: read outside [0,N] or wider than the band")
    return band, totals


def counts_all(band_f32, n_snps, W):
    out = np.zeros((n_snps + 1, 8), dtype=np.float64)
    lib().or_counts_all(_ptr(band_f32), n_snps, W, _ptr(out))
    return out


def edge_weights_at(band_f32, n_snps, W, L, snp, path, v_site="from", skip_unsym=True):
    w = np.zeros(7, dtype=np.float64)
    tw = np.zeros(1, dtype=np.float64)
    path = np.ascontiguousarray(path, dtype=np.uint8)
    mask = lib().or_edge_weights_at(_ptr(band_f32), n_snps, W, L, int(v_site == "to"), int(skip_unsym),
                                    snp, _ptr(path), _ptr(w), _ptr(tw))
    return mask, w, float(tw[0])


def generate_path(cur, orig, n_snps, W, L, v_site="from", skip_unsym=True):
    """-> (path uint8[N+1] | None, (hp_current, hp_original, min_marginal) | hole site)."""
    path = np.zeros(n_snps + 1, dtype=np.uint8)
    out = np.zeros(3, dtype=np.float64)
    rc = lib().or_generate_path(_ptr(cur), _ptr(orig), n_snps, W, L, int(v_site == "to"),
                                int(skip_unsym), _ptr(path), _ptr(out))
    if rc != 0:
        return None, rc
    return path, (float(out[0]), float(out[1]), float(out[2]))


def reweight_path(band_f32, n_snps, W, path, ratio):
    path = np.ascontiguousarray(path, dtype=np.uint8)
    return float(lib().or_reweight_path(_ptr(band_f32), n_snps, W, _ptr(path), float(ratio)))
