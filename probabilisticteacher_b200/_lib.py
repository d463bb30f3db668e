"""ctypes binding of libptb200.so (the C ABI in include/ptb200.h).

The product path has no CPU fallback: if the shared library is missing or a symbol is absent the
import fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libptb200.so")

_lib = None


class PTB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PTB200Error(
                f"{LIB_PATH} is missing: run `python -m probabilisticteacher_b200.build` "
                "(there is no CPU fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def check(code, what=""):
    if code != 0:
        raise PTB200Error(f"ptb200 call {what} failed with code {code}")


def ptr(t):
    """Device (or host) pointer of a torch tensor as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
