"""ctypes binding of libgga_b200.so (the C ABI declared in include/gga_b200.h).

There is no CPU fallback: if the library cannot be loaded the import of any op raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get('GGA_B200_LIB') or os.path.join(_HERE, '_C', 'libgga_b200.so')  # env: A/B testing of builds
_lock = threading.Lock()
_lib = None

c_float_p = ctypes.c_void_p  # device pointers travel as integers
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int


class BoxLossArgs(ctypes.Structure):
    """Mirror of `gga_box_loss_args` (include/gga_b200.h)."""
    _fields_ = [
        ('boxes', c_void_p), ('proj', c_void_p), ('proj_stride', c_int),
        ('rt', c_void_p), ('rt_stride', c_int),
        ('frame_of_box', c_void_p), ('img_hw', c_void_p), ('pcd_range', c_void_p),
        ('target', c_void_p), ('weight', c_void_p), ('weight_cols', c_int),
        ('grad_loss', c_void_p),
        ('n', c_int), ('mode', c_int), ('loss_kind', c_int), ('clamp_to_image', c_int),
        ('depth_clamp', ctypes.c_float), ('eps', ctypes.c_float), ('grad_scale', ctypes.c_float),
        ('box2d', c_void_p), ('valid', c_void_p), ('argidx', c_void_p), ('loss', c_void_p),
        ('loss_sum', c_void_p), ('loss_accum', c_void_p), ('grad_boxes', c_void_p), ('grad_box2d', c_void_p),
        ('grad_target', c_void_p),
        ('scratch', c_void_p), ('scratch_bytes', ctypes.c_size_t),
    ]


class TargetArgs(ctypes.Structure):
    """Mirror of `gga_target_args` (include/gga_b200.h)."""
    _fields_ = [
        ('labels', c_void_p), ('frame_offsets', c_void_p), ('boxes_img', c_void_p), ('lidar2img', c_void_p),
        ('pseudo', c_void_p), ('bdry', c_void_p), ('base_lidar2img', c_void_p), ('srl', c_void_p),
        ('class_task', c_void_p), ('class_cls', c_void_p), ('task_channel0', c_void_p),
        ('pseudo_dtype', ctypes.c_int32),
        ('num_frames', ctypes.c_int32), ('n_tasks', ctypes.c_int32), ('n_classes', ctypes.c_int32),
        ('n_channels', ctypes.c_int32), ('max_frame_objs', ctypes.c_int32), ('max_objs', ctypes.c_int32),
        ('fm_w', ctypes.c_int32), ('fm_h', ctypes.c_int32), ('out_size_factor', ctypes.c_int32),
        ('min_radius', ctypes.c_int32),
        ('pc_x0', ctypes.c_float), ('pc_y0', ctypes.c_float), ('voxel_x', ctypes.c_float), ('voxel_y', ctypes.c_float),
        ('gaussian_overlap', ctypes.c_double),
        ('heatmap', c_void_p), ('anno_box', c_void_p), ('ind', c_void_p), ('mask', c_void_p),
        ('anno_lidar2img', c_void_p), ('boundary_mask', c_void_p), ('src_index', c_void_p),
    ]


F32, F64 = 0, 1

# every symbol include/gga_b200.h declares, with its argument types
SIGNATURES = {
    'gga_version': ([], c_int),
    'gga_last_error': ([], ctypes.c_char_p),
    'gga_device_info': ([ctypes.POINTER(c_int)] * 3 + [ctypes.POINTER(ctypes.c_size_t)], c_int),
    'gga_pib_row_words': ([c_int], c_int),
    'gga_points_in_boxes_bits': ([c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    'gga_points_in_boxes_all': ([c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    'gga_points_in_boxes_part': ([c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    'gga_points_in_boxes_all_host': ([c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int], c_int),
    'gga_pib_hit_list': ([c_void_p, ctypes.c_int64, c_int, ctypes.c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p], c_int),
    'gga_box_project_loss': ([ctypes.POINTER(BoxLossArgs), c_void_p], c_int),
    'gga_box_project_backward': ([c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int, c_int, ctypes.c_float, c_void_p], c_int),
    'gga_loss_scratch_bytes': ([], ctypes.c_size_t),
    'gga_box2d_loss': ([c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, ctypes.c_float,
                        ctypes.c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_size_t,
                        c_void_p], c_int),
    'gga_box3d_aa_loss': ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_float, ctypes.c_float,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_size_t, c_void_p], c_int),
    'gga_match_dt_gt': ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_void_p], c_int),
    'gga_image_box_overlap_f64': ([c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p], c_int),
    'gga_points_in_convex_polygons': ([c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                       c_void_p, c_void_p], c_int),
    'gga_face_distances': ([c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p], c_int),
    'gga_pal_workspace_bytes': ([c_int, c_int], ctypes.c_size_t),
    'gga_point_box_alignment': ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 ctypes.c_size_t, c_void_p], c_int),
    'gga_pack_targets': ([ctypes.POINTER(TargetArgs), c_void_p], c_int),
    'gga_step_create': ([c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p)], c_int),
    'gga_step_destroy': ([c_void_p], c_int),
    'gga_step_device_bits': ([c_void_p, ctypes.POINTER(c_void_p)], c_int),
    'gga_step_submit_host': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                              ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                              c_void_p, c_int, c_void_p, c_void_p], c_int),
    'gga_step_wait_host': ([c_void_p, c_void_p, c_void_p], c_int),
    'gga_step_run_host_hits': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                c_void_p, c_int, c_void_p, c_void_p, c_void_p], c_int),
    'gga_step_run_host': ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                           ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                           c_void_p, c_void_p, c_void_p], c_int),
    'gga_test_sincos': ([c_void_p, ctypes.c_int64, c_void_p, c_void_p, c_void_p], c_int),
    'gga_test_box_prep': ([c_void_p, c_int, c_void_p, c_void_p], c_int),
}

# developer hooks of GGA_PROFILING builds (tools/build_prof.py); absent from the product library
DEV_SIGNATURES = {
    'gga_prof_pib': ([c_int, c_int, c_int, c_void_p], c_int),
}

PROJ_LIDAR_DIRECT, PROJ_KITTI_CAM, PROJ_CAM_CENTER, PROJ_CAM_BOTTOM = 0, 1, 2, 3
LOSS_NONE, LOSS_GIOU, LOSS_IOU_LINEAR, LOSS_IOU_SQUARE, LOSS_IOU_LOG, LOSS_L1 = 0, 1, 2, 3, 4, 5


def lib_path():
    return _LIB_PATH


def load():
    """Loads (building first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(_LIB_PATH):
            from . import build as _build
            _build.build()
        L = ctypes.CDLL(_LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.argtypes = argtypes
            fn.restype = restype
        for name, (argtypes, restype) in DEV_SIGNATURES.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.argtypes = argtypes
                fn.restype = restype
        _lib = L
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = load().gga_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'libgga_b200 {what} failed (code {rc}): {msg}')


_SCRATCH = {}


def loss_scratch(device):
    """Zeroed scratch tensor for one loss-reduction call on the CURRENT stream of `device`
    (include/gga_b200.h: caller-owned, zeroed once, not shared by concurrent calls).  Cached per
    (device, stream): calls on one stream are ordered and the kernels leave it zeroed.  Under
    CUDA-graph capture a private tensor is returned (the graph keeps it alive)."""
    import torch
    device = torch.device(device)
    nbytes = int(load().gga_loss_scratch_bytes())
    if torch.cuda.is_current_stream_capturing():
        return torch.zeros((nbytes,), dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    s = _SCRATCH.get(key)
    if s is None:
        s = torch.zeros((nbytes,), dtype=torch.uint8, device=device)
        _SCRATCH[key] = s
    return s


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return None if t is None else t.data_ptr()


def current_stream(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
