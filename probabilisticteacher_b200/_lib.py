"""ctypes binding of libptb200.so (the C ABI declared in include/ptb200.h).

The prototypes are parsed from the header, so the header is the single source of truth for the
boundary. The product path has no CPU fallback: if the shared library is missing or a declared
symbol is absent, loading fails loudly.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTB200_LIB") or os.path.join(_HERE, "libptb200.so")  # (PTB200_LIB: A/B builds in tools/)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ptb200.h")

_lib = None
_protos = None


class PTB200Error(RuntimeError):
    pass


def parse_header(path=HEADER_PATH):
    """Returns {name: [(ctype_str, param_name), ...]} for every `int ptb200_*(...)` prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(ptb200_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2)
        params = []
        for a in args.split(","):
            a = " ".join(a.split())
            mm = re.match(r"(.*?)(\w+)$", a)
            params.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = params
    return protos


def _ctype(t):
    if "*" in t:
        return ctypes.c_void_p
    return {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float}[t]


def protos():
    global _protos
    if _protos is None:
        _protos = parse_header()
    return _protos


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PTB200Error(
                f"{LIB_PATH} is missing: run `python -m probabilisticteacher_b200.build` "
                "(there is no CPU fallback for the hot path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, params in protos().items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                raise PTB200Error(f"{LIB_PATH} does not export {name} declared in {HEADER_PATH}")
            fn.restype = ctypes.c_int
            fn.argtypes = [_ctype(t) for t, _ in params]
        _lib = L
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


_keepalive = []

# kernels launched per entry point (default 1); `launch_count[0]` counts launches of OUR kernels.
KERNELS_PER_CALL = {"ptb200_nms": 8, "ptb200_rpn_match": 2, "ptb200_roi_loss_unsup": 2}
launch_count = [0]
# optional per-call profiler: set to a callable(name, args) -> context manager (bench.py uses it to
# time the dominant kernel with CUDA events on the launching stream)
profiler = [None]


_stream = [None]


def refresh_stream():
    """Re-reads torch's current CUDA stream (cached: looking it up costs ~15 us per call). Called by the
    model / trainer entry points; launches in between go to the cached stream."""
    import torch
    _stream[0] = torch.cuda.current_stream().cuda_stream
    return _stream[0]


_plans = {}


def _plan(name):
    """Per-entry-point argument plan: 0 = pass through, 1 = pointer-like, 2 = float host array,
    3 = int host array."""
    params = protos()[name]
    plan = []
    for ctype, pname in params:
        if "*" in ctype:
            if pname.endswith("_host"):
                plan.append(2 if "float" in ctype else 3)
            else:
                plan.append(1)
        else:
            plan.append(0)
    fn = getattr(lib(), name)
    _plans[name] = (fn, plan, len(plan), KERNELS_PER_CALL.get(name, 1))
    return _plans[name]


def call(name, *args):
    """Calls an entry point: tensors -> data pointers, None -> NULL, python lists for `*_host`
    parameters -> temporary C arrays; the trailing `stream` argument is filled in automatically
    when omitted. Raises PTB200Error on a non-zero return code."""
    p = _plans.get(name)
    if p is None:
        p = _plan(name)
    fn, plan, n, nk = p
    if len(args) == n - 1:
        st = _stream[0]
        if st is None:
            st = refresh_stream()
        args = args + (st,)
    elif len(args) != n:
        raise TypeError(f"{name} expects {n} arguments, got {len(args)}")
    conv = []
    for kind, a in zip(plan, args):
        if kind == 0 or a is None:
            conv.append(a)
        elif kind == 1:
            conv.append(a if isinstance(a, int) else (a.data_ptr() if hasattr(a, "data_ptr") else a))
        elif isinstance(a, (list, tuple)):
            arr = ((ctypes.c_float if kind == 2 else ctypes.c_int) * len(a))(*a)
            conv.append(ctypes.cast(arr, ctypes.c_void_p))
        else:
            conv.append(a.data_ptr() if hasattr(a, "data_ptr") else a)
    launch_count[0] += nk
    prof = profiler[0]
    if prof is not None:
        tok = prof.begin(name, args)
        rc = fn(*conv)
        prof.end(tok)
    else:
        rc = fn(*conv)
    if rc != 0:
        raise PTB200Error(f"{name} failed with code {rc}")
