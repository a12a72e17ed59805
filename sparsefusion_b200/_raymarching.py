"""Drop-in for the reference's `_raymarching` pybind module (raymarching/src/bindings.cpp:5-18).

The ten function names and positional argument orders are the reference's (raymarching.h:7-17); all
outputs are pre-allocated by the caller and written in place; every function returns None.  Point
`raymarching/raymarching.py:10`'s ``import _raymarching as _backend`` at this module with
``sys.modules['_raymarching'] = sparsefusion_b200._raymarching``.

Unlike the reference (legacy default stream, no device guard: raymarching.cu:154) every call runs
on torch's current stream of the tensors' device.
"""
from __future__ import annotations

import torch

from . import _lib as lib

f, i, b = lib.fptr, lib.iptr, lib.bptr


def _dev(t):
    if not t.is_cuda:
        raise RuntimeError('raymarching operators need CUDA tensors (the reference has no CPU path either: raymarching.py:34-35)')
    return torch.cuda.device(t.device)


def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
    with _dev(rays_o):
        lib.call('sfb_near_far_from_aabb', f(rays_o, 'rays_o'), f(rays_d, 'rays_d'), f(aabb, 'aabb'), int(N), float(min_near),
                 f(nears, 'nears'), f(fars, 'fars'), lib.stream())


def sph_from_ray(rays_o, rays_d, radius, N, coords):
    with _dev(rays_o):
        lib.call('sfb_sph_from_ray', f(rays_o, 'rays_o'), f(rays_d, 'rays_d'), float(radius), int(N), f(coords, 'coords'), lib.stream())


def morton3D(coords, N, indices):
    with _dev(coords):
        lib.call('sfb_morton3D', i(coords, 'coords'), int(N), i(indices, 'indices'), lib.stream())


def morton3D_invert(indices, N, coords):
    with _dev(indices):
        lib.call('sfb_morton3D_invert', i(indices, 'indices'), int(N), i(coords, 'coords'), lib.stream())


def packbits(grid, N, density_thresh, bitfield):
    with _dev(grid):
        lib.call('sfb_packbits', f(grid, 'grid'), int(N), float(density_thresh), b(bitfield, 'bitfield'), lib.stream())


def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises):
    with _dev(rays_o):
        lib.call('sfb_march_rays_train', f(rays_o, 'rays_o'), f(rays_d, 'rays_d'), b(grid, 'grid'), float(bound), float(dt_gamma),
                 int(max_steps), int(N), int(C), int(H), int(M), f(nears, 'nears'), f(fars, 'fars'), f(xyzs, 'xyzs'), f(dirs, 'dirs'),
                 f(deltas, 'deltas'), i(rays, 'rays'), i(counter, 'counter'), f(noises, 'noises'), lib.stream())


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image):
    with _dev(sigmas):
        lib.call('sfb_composite_rays_train_forward', f(sigmas, 'sigmas'), f(rgbs, 'rgbs'), f(deltas, 'deltas'), i(rays, 'rays'), int(M),
                 int(N), float(T_thresh), f(weights_sum, 'weights_sum'), f(depth, 'depth'), f(image, 'image'), lib.stream())


def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh,
                                  grad_sigmas, grad_rgbs):
    with _dev(sigmas):
        lib.call('sfb_composite_rays_train_backward', f(grad_weights_sum, 'grad_weights_sum'), f(grad_image, 'grad_image'),
                 f(sigmas, 'sigmas'), f(rgbs, 'rgbs'), f(deltas, 'deltas'), i(rays, 'rays'), f(weights_sum, 'weights_sum'),
                 f(image, 'image'), int(M), int(N), float(T_thresh), f(grad_sigmas, 'grad_sigmas'), f(grad_rgbs, 'grad_rgbs'), lib.stream())


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs,
               deltas, noises):
    with _dev(rays_o):
        lib.call('sfb_march_rays', int(n_alive), int(n_step), i(rays_alive, 'rays_alive'), f(rays_t, 'rays_t'), f(rays_o, 'rays_o'),
                 f(rays_d, 'rays_d'), float(bound), float(dt_gamma), int(max_steps), int(C), int(H), b(grid, 'grid'), f(nears, 'nears'),
                 f(fars, 'fars'), f(xyzs, 'xyzs'), f(dirs, 'dirs'), f(deltas, 'deltas'), f(noises, 'noises'), lib.stream())


def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    with _dev(sigmas):
        lib.call('sfb_composite_rays', int(n_alive), int(n_step), float(T_thresh), i(rays_alive, 'rays_alive'), f(rays_t, 'rays_t'),
                 f(sigmas, 'sigmas'), f(rgbs, 'rgbs'), f(deltas, 'deltas'), f(weights_sum, 'weights_sum'), f(depth, 'depth'),
                 f(image, 'image'), lib.stream())
