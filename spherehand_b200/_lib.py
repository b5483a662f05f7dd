"""ctypes binding of libspherehand_b200.so — the C-ABI boundary (include/spherehand_b200.h).

The prototypes are read from the public header itself, so the Python side can never drift from the
declared ABI.  There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, 'libspherehand_b200.so')
HEADER_PATH = os.path.join(_ROOT, 'include', 'spherehand_b200.h')

_CTYPES = {
    'void*': ctypes.c_void_p, 'const void*': ctypes.c_void_p, 'int': ctypes.c_int, 'float': ctypes.c_float,
    'long': ctypes.c_long, 'const float*': ctypes.c_void_p, 'size_t': ctypes.c_size_t, 'const char*': ctypes.c_char_p,
}


class SphereHandError(RuntimeError):
    pass


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'(const char\*|int|size_t|long)\s+(sh_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                t = a.rsplit(' ', 1)[0] if not a.endswith('*') else a
                t = t.replace(' *', '*')
                argtypes.append(t)
        protos[name] = (ret, argtypes)
    return protos


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphereHandError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                                  '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (ret, args) in _protos.items():
            fn = getattr(L, name)          # AttributeError if the library does not export a declared symbol
            fn.restype = _CTYPES[ret]
            fn.argtypes = [_CTYPES[a] for a in args]
        _lib = L
    return _lib


PROFILE = None     # set to a list to record (entry name, args, start event, end event) per call (bench.py's kernel shares)
NVTX = None        # set to a callable (entry name, args) -> range name | None to wrap calls in NVTX ranges (tools/profile_step.py)


def call(name, *args):
    """Call an int-returning entry point; raise with the library's error text on failure."""
    prof = PROFILE
    if prof is not None:
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    tag = NVTX(name, args) if NVTX is not None else None
    if tag is not None:
        import torch
        torch.cuda.nvtx.range_push(tag)
    rc = getattr(lib(), name)(*args)
    if tag is not None:
        torch.cuda.nvtx.range_pop()
    if prof is not None:
        ev1.record()
        prof.append((name, args, ev0, ev1))
    if rc != 0:
        raise SphereHandError('%s failed (%d): %s' % (name, rc, lib().sh_last_error().decode()))
