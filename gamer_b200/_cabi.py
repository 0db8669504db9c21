"""ctypes binding of libgamer_b200.so — signatures are derived from include/gamer_b200.h so they cannot drift.

The product path fails loudly when the library is missing: there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "gamer_b200.h")
LIB_PATH = os.path.join(HERE, "lib", "libgamer_b200.so")

_CTYPES = {
    "int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float,
    "gamer_stream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER) -> dict:
    """{name: (restype, [(argtype, argname), ...])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(gamer_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret:
            continue
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        decls[name] = (ret, parsed)
    return decls


def _to_ctype(t: str):
    t = t.replace("const", "").strip()
    if t.endswith("*"):
        return ctypes.c_char_p if t.replace(" ", "") == "char*" else ctypes.c_void_p
    return _CTYPES[t]


class _Lib:
    def __init__(self):
        self._lib = None
        self.decls = parse_header()

    def load(self):
        if self._lib is not None:
            return self._lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"gamer_b200: native library {LIB_PATH} is missing — run `python -m gamer_b200.build` "
                "(there is no CPU/PyTorch fallback for the hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in self.decls.items():
            fn = getattr(lib, name)  # AttributeError = header/library mismatch: fail loudly
            fn.restype = _to_ctype(ret)
            fn.argtypes = [_to_ctype(t) for t, _ in args]
        self._lib = lib
        return lib


_LIB = _Lib()


def lib():
    return _LIB.load()


def declared_symbols():
    return sorted(_LIB.decls)


class GamerError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().gamer_last_error()
        raise GamerError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


# kernels launched by one call of each entry point (for bench.py's `gpu_launches`; memsets are not counted)
LAUNCHES = {"gamer_route_perm_build": 3, "gamer_embed_sort_build": 3, "gamer_attn_fwd": 3, "gamer_attn_bwd": 3}


class Profile:
    """Optional per-entry-point device timing (CUDA events on the launching stream) + algorithmic work counters."""

    def __init__(self, only=None):
        self.records = {}      # name -> list of (start_event, end_event, flops, bytes)
        self.launches = 0
        self.only = only       # set of entry points to time (None = all); the others are only counted

    def summary(self):
        out = {}
        for name, recs in self.records.items():
            ms = sum(a.elapsed_time(b) for a, b, _, _ in recs)
            out[name] = dict(calls=len(recs), ms=ms, flops=sum(r[2] for r in recs), bytes=sum(r[3] for r in recs))
        return out


_PROFILE: Profile | None = None


def set_profile(p: Profile | None):
    global _PROFILE
    _PROFILE = p


def call(name: str, *args, work=(0, 0)):
    fn = getattr(lib(), name)
    prof = _PROFILE
    if prof is None:
        check(fn(*args), name)
        return
    import torch
    prof.launches += LAUNCHES.get(name, 1)
    if prof.records is None or (prof.only is not None and name not in prof.only):      # count-only
        check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    check(rc, name)
    prof.records.setdefault(name, []).append((e0, e1, work[0], work[1]))
