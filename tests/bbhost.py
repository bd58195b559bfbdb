"""Builds/loads tests/_bb_host.cpp (host build of the product bitboard header)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_bb_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "_bb_host.cpp")
        hdr = os.path.join(_HERE, "..", "othellozero_b200", "csrc", "oz_bitboard.cuh")
        if (not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-o", _SO, src])
        L = C.CDLL(_SO)
        u64 = C.c_uint64
        L.bbh_legal.argtypes = [u64, u64, C.c_int]
        L.bbh_legal.restype = u64
        L.bbh_flip.argtypes = [C.c_int, u64, u64]
        L.bbh_flip.restype = u64
        L.bbh_legal_compact.argtypes = [u64, u64, C.c_int]
        L.bbh_legal_compact.restype = u64
        L.bbh_flip_compact.argtypes = [C.c_int, u64, u64]
        L.bbh_flip_compact.restype = u64
        L.bbh_play.argtypes = [C.c_int, C.POINTER(u64), C.POINTER(u64), C.c_int, C.POINTER(u64)]
        L.bbh_play.restype = C.c_uint
        L.bbh_kth.argtypes = [u64, C.c_int]
        L.bbh_sm64.argtypes = [u64]
        L.bbh_sm64.restype = u64
        L.bbh_initial.argtypes = [C.c_int, C.POINTER(u64), C.POINTER(u64)]
        _lib = L
    return _lib
