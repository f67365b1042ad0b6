"""`gga_b200.compat.install()` — the advertised way into an existing mmdet3d / mmcv process
(INTEGRATION.md §A): with stand-in `mmcv.ops`, `mmdet3d.ops`, `...base_box3d` and `mmdet.models.builder`
modules in sys.modules (this container has neither package) it must replace exactly the names the
reference binds (/root/reference/mmdet3d/ops/__init__.py:12-13, base_box3d.py:7) and register the loss
modules the way `build_loss` finds them (models/builder.py:71-79).  The GPU test then calls the op
through the patched module, the way `BaseInstance3DBoxes.points_in_boxes_all` does (base_box3d.py:566)."""
import sys
import types

import numpy as np
import pytest
import torch

FAKE = ('mmcv', 'mmcv.ops', 'mmcv.ops.points_in_boxes', 'mmdet3d', 'mmdet3d.ops', 'mmdet3d.core', 'mmdet3d.core.bbox',
        'mmdet3d.core.bbox.structures', 'mmdet3d.core.bbox.structures.base_box3d', 'mmdet', 'mmdet.models',
        'mmdet.models.builder')


class _Registry:
    def __init__(self):
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        assert module is not None
        key = name or module.__name__
        if key in self.module_dict and not force:
            raise KeyError(key)
        self.module_dict[key] = module
        return module


@pytest.fixture
def fake_env():
    saved = {k: sys.modules.get(k) for k in FAKE}

    def stock(*a, **k):
        raise RuntimeError('the stock mmcv op was called')
    for name in FAKE:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    for name in ('mmcv.ops', 'mmcv.ops.points_in_boxes', 'mmdet3d.ops'):
        for fn in ('points_in_boxes_all', 'points_in_boxes_part', 'points_in_boxes_cpu'):
            setattr(sys.modules[name], fn, stock)
    bb = sys.modules['mmdet3d.core.bbox.structures.base_box3d']
    bb.points_in_boxes_all, bb.points_in_boxes_part = stock, stock       # `from mmcv.ops import ...` (base_box3d.py:7)
    sys.modules['mmcv.ops'].box_iou_rotated = stock                      # must stay untouched
    sys.modules['mmdet.models.builder'].LOSSES = _Registry()
    yield stock
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_install_replaces_the_bound_names_and_registers_losses(fake_env):
    from gga_b200 import compat, losses, ops
    done = compat.install()
    for name in ('mmcv.ops', 'mmcv.ops.points_in_boxes', 'mmdet3d.ops'):
        m = sys.modules[name]
        assert m.points_in_boxes_all is ops.points_in_boxes_all
        assert m.points_in_boxes_part is ops.points_in_boxes_part
        assert m.points_in_boxes_cpu is ops.points_in_boxes_cpu
    bb = sys.modules['mmdet3d.core.bbox.structures.base_box3d']
    assert bb.points_in_boxes_all is ops.points_in_boxes_all and bb.points_in_boxes_part is ops.points_in_boxes_part
    assert not hasattr(bb, 'points_in_boxes_cpu')                        # only names that were bound are replaced
    assert sys.modules['mmcv.ops'].box_iou_rotated is fake_env
    reg = sys.modules['mmdet.models.builder'].LOSSES.module_dict
    assert reg['ProjectedGIoULoss'] is losses.ProjectedGIoULoss and reg['ProjectedL1Loss'] is losses.ProjectedL1Loss
    assert reg['ProjectedIoULoss'] is losses.ProjectedIoULoss and reg['AxisAlignedIoULoss'] is losses.AxisAlignedIoULoss
    assert 'mmcv.ops.points_in_boxes_all' in done and any('LOSSES.ProjectedGIoULoss' in d for d in done)
    # idempotent
    assert set(compat.install()) == set(done)


def test_install_without_the_packages_is_a_no_op():
    from gga_b200 import compat
    if any(k in sys.modules for k in ('mmcv', 'mmdet3d', 'mmdet')):
        pytest.skip('a real mmcv / mmdet is importable here')
    assert compat.install() == []


@pytest.mark.gpu
def test_patched_module_serves_the_box_class_call(fake_env):
    """The call `points_in_boxes_all(points_clone, boxes)` of base_box3d.py:559-568 through the patched name."""
    from gga_b200 import compat, synth
    from oracle import membership as om
    compat.install()
    bb = sys.modules['mmdet3d.core.bbox.structures.base_box3d']
    f = synth.make_frame(1, 5, N=4000)
    points = torch.from_numpy(f['points']).cuda()
    boxes = torch.from_numpy(f['boxes']).cuda()
    pc = points.clone()[..., :3]                                          # :559
    pc = pc.unsqueeze(0)                                                  # :561-562
    out = bb.points_in_boxes_all(pc, boxes.unsqueeze(0).to(pc.device))    # :565-566
    ref = om.points_in_boxes_all_np(f['points'], f['boxes'], 8)
    assert np.array_equal(out.squeeze(0).cpu().numpy(), ref)
    part = bb.points_in_boxes_part(pc, boxes.unsqueeze(0)).squeeze(0).cpu().numpy()
    assert np.array_equal(part, np.where(ref.any(1), ref.argmax(1), -1))
