"""The C-ABI library loads on a CPU-only box and exports every symbol include/sparsefusion_b200.h declares."""
import ctypes
import os
import subprocess

import pytest


def test_library_exports_every_declared_symbol():
    from sparsefusion_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in protos if not hasattr(lib, name)]
    assert not missing, f'declared in the header but not exported: {missing}'
    exported = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    extra = [l.split()[-1] for l in exported.splitlines() if ' T sfb_' in l and l.split()[-1] not in protos]
    assert not extra, f'exported but not declared in the header: {extra}'


def test_load_sets_signatures_and_reports_errors():
    from sparsefusion_b200 import _lib
    lib = _lib.load()
    assert lib.sfb_abi_version() == 3
    # argument validation happens before any CUDA call, so it is testable without a GPU
    with pytest.raises(RuntimeError, match='null pointer'):
        _lib.call('sfb_near_far_from_aabb', None, None, None, 4, 0.1, None, None, None)
    with pytest.raises(RuntimeError, match='D must be 1..5'):
        _lib.call('sfb_grid_encode_forward', 16, 16, 16, 16, 1, 7, 2, 16, 0.5, 16, None, 1, 0, None)
    assert _lib.load().sfb_conv_weight_k(260, 3, 3) == 9 * 288


def test_host_wrappers_refuse_cpu_tensors():
    """no CPU fallback: the product path must fail loudly without CUDA"""
    import torch
    from sparsefusion_b200 import _gridencoder, _raymarching
    with pytest.raises(RuntimeError):
        _raymarching.near_far_from_aabb(torch.zeros(4, 3), torch.ones(4, 3), torch.zeros(6), 4, 0.1, torch.zeros(4), torch.zeros(4))
    with pytest.raises(RuntimeError):
        _gridencoder.grid_encode_forward(torch.zeros(4, 3), torch.zeros(8, 2), torch.zeros(3, dtype=torch.int32), torch.zeros(2, 4, 2), 4, 3, 2, 2,
                                         0.5, 16, None, 1, False)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, 'sparsefusion_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'oracle/' not in text.replace('oracle/_ref', ''), f
