"""Drop-in for the reference's `_gridencoder` pybind module (external/gridencoder/src/bindings.cpp:5-8).

Same two function names, same positional argument order, same conventions (pre-allocated
outputs, in-place writes, returns None) -- so `external/gridencoder/grid.py:10`'s
``import _gridencoder as _backend`` can be pointed at this module unchanged
(``sys.modules['_gridencoder'] = sparsefusion_b200._gridencoder``).
Each call forwards device pointers to the C ABI on torch's current stream.
"""
from __future__ import annotations

import torch

from . import _lib as lib


def _check(inputs, embeddings, offsets):
    # mirrors CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_INT / CHECK_IS_FLOATING (gridencoder.cu:425-441)
    for name, t in (('inputs', inputs), ('embeddings', embeddings), ('offsets', offsets)):
        if not t.is_cuda:
            raise RuntimeError(f'{name} must be a CUDA tensor')
        if not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous tensor')
    if offsets.dtype != torch.int32:
        raise RuntimeError('offsets must be an int tensor')
    if embeddings.dtype != torch.float32 or inputs.dtype != torch.float32:
        raise RuntimeError('sparsefusion_b200 grid encoder is fp32 only (the reference path never runs half: SURVEY.md §2.3)')


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners):
    _check(inputs, embeddings, offsets)
    with torch.cuda.device(inputs.device):
        lib.call('sfb_grid_encode_forward', lib.fptr(inputs, 'inputs'), lib.fptr(embeddings, 'embeddings'), lib.iptr(offsets, 'offsets'),
               lib.fptr(outputs, 'outputs'), int(B), int(D), int(C), int(L), float(S), int(H), lib.fptr(dy_dx, 'dy_dx'),
               int(gridtype), int(bool(align_corners)), lib.stream())


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype, align_corners):
    _check(inputs, embeddings, offsets)
    with torch.cuda.device(inputs.device):
        lib.call('sfb_grid_encode_backward', lib.fptr(grad, 'grad'), lib.fptr(inputs, 'inputs'), lib.fptr(embeddings, 'embeddings'),
               lib.iptr(offsets, 'offsets'), lib.fptr(grad_embeddings, 'grad_embeddings'), int(B), int(D), int(C), int(L), float(S),
               int(H), lib.fptr(dy_dx, 'dy_dx'), lib.fptr(grad_inputs, 'grad_inputs'), int(gridtype), int(bool(align_corners)), lib.stream())
