"""Drop-in installation into an existing mmdet3d / mmcv environment.

``install()`` makes the reference pick up this library without touching its sources:
* ``mmcv.ops.points_in_boxes_{all,part,cpu}`` (and the names already bound inside
  ``mmdet3d.core.bbox.structures.base_box3d`` / ``mmdet3d.ops``) are replaced;
* the loss modules are registered in mmdet's ``LOSSES`` registry under
  ``ProjectedGIoULoss / ProjectedIoULoss / ProjectedL1Loss`` so configs can select them with
  ``loss_consistency=dict(type='ProjectedGIoULoss', loss_weight=1.0)`` (cf. ``pgd_head.py:72``).
Every step is skipped silently when the target package is absent (this container has neither
mmcv nor mmdet); the CUDA library itself must load — there is no fallback.
"""
import sys

from . import losses, ops


def install(patch_mmcv=True, register_losses=True):
    done = []
    if patch_mmcv:
        for modname in ('mmcv.ops', 'mmcv.ops.points_in_boxes', 'mmdet3d.ops',
                        'mmdet3d.core.bbox.structures.base_box3d'):
            m = sys.modules.get(modname)
            if m is None:
                try:
                    m = __import__(modname, fromlist=['_'])
                except Exception:
                    continue
            for name in ('points_in_boxes_all', 'points_in_boxes_part', 'points_in_boxes_cpu'):
                if hasattr(m, name):
                    setattr(m, name, getattr(ops, name))
                    done.append(f'{modname}.{name}')
    if register_losses:
        for regmod in ('mmdet3d.models.builder', 'mmdet.models.builder'):
            try:
                reg = __import__(regmod, fromlist=['LOSSES']).LOSSES
            except Exception:
                continue
            for cls in (losses.ProjectedGIoULoss, losses.ProjectedIoULoss, losses.ProjectedL1Loss,
                        losses.AxisAlignedIoULoss):
                try:
                    reg.register_module(module=cls, force=True)
                    done.append(f'{regmod}.LOSSES.{cls.__name__}')
                except Exception:
                    pass
            break
    return done
