"""ctypes binding to libsparsefusion_b200.so (C ABI: include/sparsefusion_b200.h).

There is deliberately NO fallback: if the library is missing, or a kernel is asked to run on a
non-CUDA tensor, this module raises.  A product path that silently ran on the CPU (or on the
test oracle) would void every parity claim.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsparsefusion_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'sparsefusion_b200.h')

_CTYPE = (
    (re.compile(r'\*'), ctypes.c_void_p),
    (re.compile(r'\buint64_t\b|\bsize_t\b'), ctypes.c_uint64),
    (re.compile(r'\bint64_t\b'), ctypes.c_int64),
    (re.compile(r'\buint32_t\b'), ctypes.c_uint32),
    (re.compile(r'\bint32_t\b|\bint\b'), ctypes.c_int),
    (re.compile(r'\bfloat\b'), ctypes.c_float),
    (re.compile(r'\bdouble\b'), ctypes.c_double),
)


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[str, List[str]]]:
    """{symbol: (return type, [argument declarations])} for every prototype in the C-ABI header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = re.sub(r'//[^\n]*', ' ', text)
    protos = {}
    for m in re.finditer(r'\b(const\s+char\s*\*|int|void|uint32_t|uint64_t|int64_t)\s+(sfb_\w+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = [a.strip() for a in re.sub(r'\s+', ' ', args).split(',')]
        if args == ['void'] or args == ['']:
            args = []
        protos[name] = (ret.strip(), args)
    return protos


def _argtype(decl: str):
    for rx, ct in _CTYPE:
        if rx.search(decl):
            return ct
    raise ValueError(f'cannot map C declaration {decl!r}')


_lib: Optional[ctypes.CDLL] = None
_protos: Dict[str, Tuple[str, List[str]]] = {}


def load() -> ctypes.CDLL:
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -m sparsefusion_b200.build` '
                           '(or __graft_entry__.build()).  There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (ret, args) in _protos.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.argtypes = [_argtype(a) for a in args]
        fn.restype = {'void': None, 'int': ctypes.c_int, 'uint32_t': ctypes.c_uint32, 'uint64_t': ctypes.c_uint64,
                      'int64_t': ctypes.c_int64}.get(ret, ctypes.c_char_p if 'char' in ret else ctypes.c_int)
    if lib.sfb_abi_version() != 3:
        raise RuntimeError('libsparsefusion_b200.so ABI version mismatch')
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke an `int sfb_*` entry point; raise RuntimeError(sfb_last_error()) on failure
    (the reference's operators raise RuntimeError through TORCH_CHECK, gridencoder.cu:425-441)."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name}: {lib.sfb_last_error().decode()}')


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None, name: str = 'tensor') -> Optional[int]:
    """device pointer of a contiguous CUDA tensor (None passes through as NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be a contiguous tensor')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'{name} must be a {dtype} tensor (got {t.dtype})')
    return t.data_ptr()


def fptr(t, name='tensor'):
    return ptr(t, torch.float32, name)


def iptr(t, name='tensor'):
    return ptr(t, torch.int32, name)


def bptr(t, name='tensor'):
    return ptr(t, torch.uint8, name)


def stream() -> int:
    """the CURRENT torch stream of the current device (the reference launches on the legacy default
    stream with no device guard, raymarching.cu:154; taking torch's current stream fixes that)"""
    return torch.cuda.current_stream().cuda_stream
