"""TEST INFRASTRUCTURE — loads the reference's own Python files read-only.

Only usable where ``/root/reference`` exists (the build container).  It is used by
``oracle/gen_golden.py`` to generate the committed fixtures under ``tests/golden/``
and by the container-only validation tests; nothing on the GPU box imports it
(``/root/reference`` does not exist there), and nothing in ``gga_b200/`` imports it.

``import mmdet3d`` fails here (no mmcv / mmdet wheels, no network), so the geometry
modules of the reference are loaded *by path* under their real module names, with
namespace stubs for the packages in between and a stub ``mmcv.ops`` exposing the three
names ``base_box3d.py:7`` imports.  The membership ops injected into that stub are
chosen by the caller: the CPU oracle (validation here) or the CUDA op (GPU box).
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get('GGA_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'mmdet3d/core/bbox/structures/utils.py'))


def _ns(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    return m


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


_CACHE = {}


def load_reference(points_in_boxes_all=None, points_in_boxes_part=None):
    """Returns a namespace with the reference's own geometry callables.

    The two arguments become ``mmcv.ops.points_in_boxes_all/_part`` as seen by the
    reference's box classes (``base_box3d.py:7,534,566``).
    """
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    if 'ns' in _CACHE:
        ops = sys.modules['mmcv.ops']
        bb = sys.modules['mmdet3d.core.bbox.structures.base_box3d']
        if points_in_boxes_all is not None:
            ops.points_in_boxes_all = points_in_boxes_all
            bb.points_in_boxes_all = points_in_boxes_all
        if points_in_boxes_part is not None:
            ops.points_in_boxes_part = points_in_boxes_part
            bb.points_in_boxes_part = points_in_boxes_part
        return _CACHE['ns']

    for n in ['mmdet3d', 'mmdet3d.core', 'mmdet3d.core.utils', 'mmdet3d.core.bbox',
              'mmdet3d.core.bbox.structures', 'mmdet3d.core.points', 'mmcv', 'mmcv.ops',
              'mmdet3d.core.evaluation', 'mmdet3d.core.evaluation.kitti_utils',
              'mmdet3d.core.bbox.iou_calculators', 'mmdet', 'mmdet.core', 'mmdet.core.bbox',
              'mmdet.core.bbox.iou_calculators', 'mmdet.core.bbox.iou_calculators.builder']:
        _ns(n)

    def _missing(*a, **k):
        raise RuntimeError('membership op not injected into the reference loader')

    ops = sys.modules['mmcv.ops']
    ops.box_iou_rotated = None
    ops.points_in_boxes_all = points_in_boxes_all or _missing
    ops.points_in_boxes_part = points_in_boxes_part or _missing

    ac = _load('mmdet3d.core.utils.array_converter', 'mmdet3d/core/utils/array_converter.py')
    sys.modules['mmdet3d.core.utils'].array_converter = ac.array_converter
    sys.modules['mmdet3d.core.utils'].ArrayConverter = ac.ArrayConverter

    S = 'mmdet3d/core/bbox/structures/'
    u = _load('mmdet3d.core.bbox.structures.utils', S + 'utils.py')
    st = sys.modules['mmdet3d.core.bbox.structures']
    for k in ('limit_period', 'points_cam2img', 'rotation_3d_in_axis', 'points_img2cam',
              'xywhr2xyxyr', 'get_box_type', 'mono_cam_box2vis', 'get_proj_mat_by_coord_type',
              'yaw2local'):
        setattr(st, k, getattr(u, k))

    P = 'mmdet3d/core/points/'
    bp = _load('mmdet3d.core.points.base_points', P + 'base_points.py')
    cp = _load('mmdet3d.core.points.cam_points', P + 'cam_points.py')
    dp = _load('mmdet3d.core.points.depth_points', P + 'depth_points.py')
    lp = _load('mmdet3d.core.points.lidar_points', P + 'lidar_points.py')
    pts = sys.modules['mmdet3d.core.points']
    pts.BasePoints, pts.CameraPoints = bp.BasePoints, cp.CameraPoints
    pts.DepthPoints, pts.LiDARPoints = dp.DepthPoints, lp.LiDARPoints

    bb = _load('mmdet3d.core.bbox.structures.base_box3d', S + 'base_box3d.py')
    lb = _load('mmdet3d.core.bbox.structures.lidar_box3d', S + 'lidar_box3d.py')
    cb = _load('mmdet3d.core.bbox.structures.cam_box3d', S + 'cam_box3d.py')
    db = _load('mmdet3d.core.bbox.structures.depth_box3d', S + 'depth_box3d.py')
    bm = _load('mmdet3d.core.bbox.structures.box_3d_mode', S + 'box_3d_mode.py')
    cm = _load('mmdet3d.core.bbox.structures.coord_3d_mode', S + 'coord_3d_mode.py')
    for mod in (st, sys.modules['mmdet3d.core.bbox']):
        mod.BaseInstance3DBoxes = bb.BaseInstance3DBoxes
        mod.LiDARInstance3DBoxes = lb.LiDARInstance3DBoxes
        mod.CameraInstance3DBoxes = cb.CameraInstance3DBoxes
        mod.DepthInstance3DBoxes = db.DepthInstance3DBoxes
        mod.Box3DMode = bm.Box3DMode
        mod.Coord3DMode = cm.Coord3DMode
        for k in ('limit_period', 'points_cam2img', 'rotation_3d_in_axis', 'get_box_type'):
            setattr(mod, k, getattr(u, k))
    sys.modules['mmdet3d.core.bbox'].structures = st

    np_ops = _load('mmdet3d.core.bbox.box_np_ops', 'mmdet3d/core/bbox/box_np_ops.py')
    sys.modules['mmdet3d.core.bbox'].box_np_ops = np_ops
    ev = _load('mmdet3d.core.evaluation.kitti_utils.eval',
               'mmdet3d/core/evaluation/kitti_utils/eval.py')

    # iou3d_calculator.py:3-5 pulls two mmdet names at import time; only the in-file
    # axis_aligned_bbox_overlaps_3d (:210-329) is used from it.
    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c
    sys.modules['mmdet.core.bbox'].bbox_overlaps = None
    sys.modules['mmdet.core.bbox.iou_calculators.builder'].IOU_CALCULATORS = _Reg()
    iou3d = _load('mmdet3d.core.bbox.iou_calculators.iou3d_calculator',
                  'mmdet3d/core/bbox/iou_calculators/iou3d_calculator.py')

    ns = types.SimpleNamespace(
        utils=u, limit_period=u.limit_period, points_cam2img=u.points_cam2img,
        rotation_3d_in_axis=u.rotation_3d_in_axis, points_img2cam=u.points_img2cam,
        BaseInstance3DBoxes=bb.BaseInstance3DBoxes, LiDARInstance3DBoxes=lb.LiDARInstance3DBoxes,
        CameraInstance3DBoxes=cb.CameraInstance3DBoxes, DepthInstance3DBoxes=db.DepthInstance3DBoxes,
        Box3DMode=bm.Box3DMode, Coord3DMode=cm.Coord3DMode,
        DepthPoints=dp.DepthPoints, LiDARPoints=lp.LiDARPoints, CameraPoints=cp.CameraPoints,
        box_np_ops=np_ops, kitti_eval=ev, image_box_overlap=ev.image_box_overlap,
        axis_aligned_bbox_overlaps_3d=iou3d.axis_aligned_bbox_overlaps_3d,
        mmcv_ops=ops)
    _CACHE['ns'] = ns
    return ns


def expose_for_reference_tests():
    """Completes the module tree so that the reference's own test files
    (tests/test_utils/test_box3d.py, ...) can be imported: `mmdet3d.core.bbox.transforms`
    (bbox3d2roi, bbox3d_mapping_back; transforms.py needs only torch) and the public names the
    tests import from `mmdet3d.core.bbox` / `mmdet3d.core.bbox.structures.utils`."""
    load_reference()
    tr = _load('mmdet3d.core.bbox.transforms', 'mmdet3d/core/bbox/transforms.py')
    core_bbox = sys.modules['mmdet3d.core.bbox']
    for k in ('bbox3d2roi', 'bbox3d_mapping_back', 'bbox3d2result'):
        if hasattr(tr, k):
            setattr(core_bbox, k, getattr(tr, k))
    return core_bbox


def load_head_functions(train_cfg, norm_bbox=True):
    """The GGA head methods of the reference, extracted from its source file and bound to a
    stand-in ``self`` (the module itself cannot be imported: its header needs mmcv.cnn and the
    mmdet registries, ``centerpoint_head_gga.py:5-15``).  Returns a namespace with the
    reference's OWN ``GGA_calculate_rotation`` (:167-182), ``get_distance_single`` (:184-239),
    ``get_distance_bev`` (:241-248) and ``get_prediction_single`` (:250-341), executing the
    reference source text unmodified."""
    import ast
    import numpy as np
    import torch
    ns = load_reference()
    path = os.path.join(REF_ROOT, 'mmdet3d/models/dense_heads/centerpoint_head_gga.py')
    tree = ast.parse(open(path).read(), filename=path)
    wanted = {'GGA_calculate_rotation', 'get_distance_single', 'get_distance_bev', 'get_prediction_single'}
    funcs = [n for cls in tree.body if isinstance(cls, ast.ClassDef) and cls.name == 'CenterHead_GGA'
             for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {f.name for f in funcs} == wanted, [f.name for f in funcs]

    def multi_apply(func, *args, **kwargs):   # mmdet.core.multi_apply [un-vendored], restated
        from functools import partial
        pfunc = partial(func, **kwargs) if kwargs else func
        return tuple(map(list, zip(*map(pfunc, *args))))

    glb = {'torch': torch, 'np': np, 'multi_apply': multi_apply, 'rotation_3d_in_axis': ns.rotation_3d_in_axis,
           '__builtins__': __builtins__}
    mod = ast.Module(body=funcs, type_ignores=[])
    exec(compile(mod, path, 'exec'), glb)
    self = types.SimpleNamespace(train_cfg=train_cfg, norm_bbox=norm_bbox)
    out = types.SimpleNamespace()
    for name in wanted:
        setattr(self, name, types.MethodType(glb[name], self))
        setattr(out, name, getattr(self, name))
    return out


def load_target_functions(train_cfg, class_names):
    """The reference's OWN ``CenterHead_GGA.get_targets_single`` (centerpoint_head_gga.py:401-627),
    its source text executed unmodified against a stand-in ``self``, with the reference's own
    ``draw_heatmap_gaussian`` / ``gaussian_radius`` (mmdet3d/core/utils/gaussian.py, loaded from
    its file: it only needs numpy and torch).  Returns the bound method."""
    import ast
    import importlib.util
    import numpy as np
    import torch
    gpath = os.path.join(REF_ROOT, 'mmdet3d/core/utils/gaussian.py')
    spec = importlib.util.spec_from_file_location('_ref_gaussian', gpath)
    gmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gmod)
    path = os.path.join(REF_ROOT, 'mmdet3d/models/dense_heads/centerpoint_head_gga.py')
    tree = ast.parse(open(path).read(), filename=path)
    funcs = [n for cls in tree.body if isinstance(cls, ast.ClassDef) and cls.name == 'CenterHead_GGA'
             for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'get_targets_single']
    assert len(funcs) == 1
    glb = {'torch': torch, 'np': np, 'draw_heatmap_gaussian': gmod.draw_heatmap_gaussian,
           'gaussian_radius': gmod.gaussian_radius, '__builtins__': __builtins__}
    exec(compile(ast.Module(body=funcs, type_ignores=[]), path, 'exec'), glb)
    self = types.SimpleNamespace(train_cfg=train_cfg, class_names=class_names, task_heads=[None] * len(class_names),
                                 with_velocity=False)
    return types.MethodType(glb['get_targets_single'], self)


def load_format_functions(data_infos, pcd_limit_range):
    """The reference's OWN ``KittiDataset_GGA_match.bbox2result_kitti`` (:458-571) and
    ``convert_valid_bboxes`` (:685-765), plus ``pseudo_label_matching_kitti``
    (tools/utils_pseudo_labels_gga.py:17-88): source text executed unmodified against a stand-in
    ``self`` (the dataset module needs the mmdet registries).  ``mmcv`` is a stub with
    ``mkdir_or_exist`` / ``track_iter_progress`` / ``dump`` (dump keeps the object in
    ``ns.dumped``).  Returns a namespace (bbox2result_kitti, pseudo_label_matching_kitti, dumped)."""
    import ast
    import copy
    import numpy as np
    import torch
    ns = load_reference()
    out = types.SimpleNamespace(dumped=[])
    mm = types.SimpleNamespace(mkdir_or_exist=lambda d: os.makedirs(d, exist_ok=True),
                               track_iter_progress=lambda it: it,
                               dump=lambda obj, path: out.dumped.append((path, obj)))
    path = os.path.join(REF_ROOT, 'mmdet3d/datasets/kitti_dataset_GGA_match.py')
    tree = ast.parse(open(path).read(), filename=path)
    wanted = {'bbox2result_kitti', 'convert_valid_bboxes'}
    funcs = [n for cls in tree.body if isinstance(cls, ast.ClassDef) and cls.name == 'KittiDataset_GGA_match'
             for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {f.name for f in funcs} == wanted
    glb = {'torch': torch, 'np': np, 'mmcv': mm, 'Box3DMode': ns.Box3DMode, 'points_cam2img': ns.points_cam2img,
           '__builtins__': __builtins__}
    exec(compile(ast.Module(body=funcs, type_ignores=[]), path, 'exec'), glb)
    self = types.SimpleNamespace(data_infos=data_infos, pcd_limit_range=pcd_limit_range)
    for name in wanted:
        setattr(self, name, types.MethodType(glb[name], self))
    out.bbox2result_kitti = self.bbox2result_kitti
    path2 = os.path.join(REF_ROOT, 'tools/utils_pseudo_labels_gga.py')
    tree2 = ast.parse(open(path2).read(), filename=path2)
    funcs2 = [n for n in tree2.body if isinstance(n, ast.FunctionDef)
              and n.name in ('drop_arrays_by_name', 'pseudo_label_matching_kitti')]
    glb2 = {'np': np, 'copy': copy, 'mmcv': mm, 'get_split_parts': ns.kitti_eval.get_split_parts,
            'calculate_iou_partly': ns.kitti_eval.calculate_iou_partly, '__builtins__': __builtins__}
    exec(compile(ast.Module(body=funcs2, type_ignores=[]), path2, 'exec'), glb2)
    out.pseudo_label_matching_kitti = glb2['pseudo_label_matching_kitti']
    out.LiDARInstance3DBoxes = ns.LiDARInstance3DBoxes
    return out
