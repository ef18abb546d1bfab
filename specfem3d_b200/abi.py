"""ctypes binding of ``libspecfem3d_b200.so`` that calls the entry points the way the Fortran solver does:
every argument by reference, gfortran-mangled names (``name_``), ``integer(kind=8)`` handle.

Python scalars are boxed according to the C prototype recorded in ``abi_symbols.json`` (generated from the
reference's stub file by tools/gen_abi.py): ``int*`` <- int/bool, ``realw*`` <- float (as float32), ``double*`` <-
float, ``long*`` <- :class:`ctypes.c_long`.  numpy arrays are passed as their data pointer (they must be
C-contiguous and of the matching dtype: the caller owns them, exactly like Fortran owns its arrays).

This module is the *product's* host-side boundary: it refuses to work without the CUDA library (no CPU fallback).
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Any

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspecfem3d_b200.so")

_CTYPE = {"int": ctypes.c_int, "realw": ctypes.c_float, "field": ctypes.c_float, "double": ctypes.c_double,
          "long": ctypes.c_long, "float": ctypes.c_float}
_NPTYPE = {"int": np.int32, "realw": np.float32, "field": np.float32, "double": np.float64, "long": np.int64,
           "float": np.float32}


class DeviceLibraryMissing(RuntimeError):
    pass


def _base_type(arg_decl: str) -> str:
    d = arg_decl.replace("const", " ").replace("*", " * ").split()
    return d[0]


class DeviceLibrary:
    """Loaded ``libspecfem3d_b200.so`` with Fortran-style call wrappers: ``lib.update_displacement_cuda(mesh, dt, ...)``."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise DeviceLibraryMissing(
                f"{path} not found: build it with `python -m specfem3d_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback for the device path.")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        with open(os.path.join(_HERE, "abi_symbols.json")) as f:
            self.symbols = json.load(f)
        self._keep = []

    def has(self, name: str) -> bool:
        try:
            getattr(self.cdll, name + "_")
            return True
        except AttributeError:
            return False

    def __getattr__(self, name: str):
        if name.startswith("_") or name in ("cdll", "symbols", "path"):
            raise AttributeError(name)
        if name not in self.symbols:
            # extension entry points (not in the reference contract): same by-reference convention for ctypes scalars
            raw = getattr(self.cdll, name + "_")
            raw.restype = None

            def ext_call(*args: Any):
                boxed = []
                for a in args:
                    if isinstance(a, np.ndarray):
                        boxed.append(a.ctypes.data_as(ctypes.c_void_p))
                    elif isinstance(a, ctypes._SimpleCData):
                        boxed.append(ctypes.byref(a))
                    elif isinstance(a, (bytes, ctypes.Array)) or a is None:
                        boxed.append(a)
                    elif isinstance(a, int):
                        boxed.append(ctypes.byref(ctypes.c_int(a)))
                    else:
                        raise TypeError(f"{name}: cannot pass {type(a)} to an extension entry point")
                return raw(*boxed)

            ext_call.__name__ = name
            return ext_call
        fn = getattr(self.cdll, name + "_")
        fn.restype = None
        decls = self.symbols[name]["args"]

        def call(*args: Any):
            if len(args) != len(decls):
                raise TypeError(f"{name} expects {len(decls)} arguments ({decls}), got {len(args)}")
            boxed = []
            for a, d in zip(args, decls):
                bt = _base_type(d)
                if isinstance(a, np.ndarray):
                    want = _NPTYPE.get(bt)
                    if want is not None and a.dtype != want:
                        raise TypeError(f"{name}: argument '{d}' needs dtype {want.__name__}, got {a.dtype}")
                    if not a.flags["C_CONTIGUOUS"] and not a.flags["F_CONTIGUOUS"]:
                        raise TypeError(f"{name}: argument '{d}' must be contiguous")
                    boxed.append(a.ctypes.data_as(ctypes.c_void_p))
                elif isinstance(a, ctypes._SimpleCData):
                    boxed.append(ctypes.byref(a))
                elif a is None:
                    boxed.append(None)
                else:
                    ct = _CTYPE[bt]
                    boxed.append(ctypes.byref(ct(a)))
            return fn(*boxed)

        call.__name__ = name
        return call


_LIB = None


def load() -> DeviceLibrary:
    global _LIB
    if _LIB is None:
        _LIB = DeviceLibrary()
    return _LIB
