"""CPU tests (no GPU, no compute calls): the C-ABI library builds, loads and exports every
symbol include/gga_b200.h declares; host-side helpers behave like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'gga_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gga_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    from gga_b200 import _lib, build
    path = build.build()
    assert os.path.isfile(path)
    L = ctypes.CDLL(path)
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/gga_b200.h but not exported'
    # the ctypes table of the Python host side covers the same set
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)


def test_pure_abi_calls_without_a_device():
    from gga_b200 import _lib
    L = _lib.load()
    assert L.gga_version() >= 100
    assert [L.gga_pib_row_words(t) for t in (0, 1, 32, 33, 64, 65, 128, 129, 256, 257, 1024)] == \
        [0, 1, 1, 2, 2, 4, 4, 8, 8, 16, 32]
    assert 4096 < L.gga_loss_scratch_bytes() <= 8192
    # argument validation happens before any CUDA call
    assert L.gga_points_in_boxes_bits(None, 2, None, None, 1, 1, 1, None) == -1
    assert b'pts_stride' in L.gga_last_error()
    assert L.gga_points_in_boxes_all(None, 3, None, None, -1, 1, 1, None) == -1
    assert L.gga_box_project_loss(None, None) == -1


def test_ops_reject_cpu_tensors_instead_of_falling_back():
    import gga_b200 as G
    p, b = torch.zeros((1, 4, 3)), torch.zeros((1, 2, 7))
    with pytest.raises(AssertionError):
        G.points_in_boxes_all(p, b)
    with pytest.raises(AssertionError):
        G.box3d_project(torch.zeros((2, 7)), torch.eye(4))
    with pytest.raises(AssertionError):
        G.box2d_loss(torch.zeros((2, 4)), torch.zeros((2, 4)))
    with pytest.raises(AssertionError):     # the mmcv shape asserts come first
        G.points_in_boxes_all(torch.zeros((1, 4, 4)), b)


def test_pad_proj_matches_points_cam2img_padding():
    import gga_b200 as G
    m = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    p = G.pad_proj(m)
    assert p.shape == (4, 4) and torch.equal(p[:3], m) and torch.equal(p[3], torch.tensor([0., 0, 0, 1]))
    assert torch.equal(G.pad_proj(torch.eye(3))[:3, :3], torch.eye(3))
    with pytest.raises(AssertionError):
        G.pad_proj(torch.zeros((2, 4)))


def test_fix_matched_dims_like_reference():
    from gga_b200.matching import fix_matched_dims
    dims = np.array([[1.0, 2.0, 3.0], [3.0, 2.0, 1.0]])
    ry = np.array([0.1, 0.2])
    d, r = fix_matched_dims(dims, ry)           # tools/utils_pseudo_labels_gga.py:74-78
    assert np.allclose(d, [[3.0, 2.0, 1.0], [3.0, 2.0, 1.0]])
    assert np.allclose(r, [0.1 + np.pi / 2, 0.2])


def test_synthetic_frames_are_deterministic_and_shaped():
    from gga_b200 import synth
    a, b = synth.make_frame(2, 5, N=2000), synth.make_frame(2, 5, N=2000)
    assert all(np.array_equal(a[k], b[k]) for k in ('points', 'boxes', 'target'))
    assert a['points'].shape == (2000, 4) and a['boxes'].shape == (256, 7) and a['target'].shape == (256, 4)
    bt = synth.make_batch(3, 0, 2, N=500)
    assert bt['points'].shape == (2, 500, 4) and bt['lidar2img'].shape == (2, 512, 4, 4)
