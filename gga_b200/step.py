"""One training-shaped step of the geometry hot path over a batch of frames:

    membership masks (bit-packed)  +  3D box -> 2D box projection  +  IoU/GIoU/L1 loss
    forward AND backward to the box parameters,

i.e. what a GGA head does per iteration with ``points_in_boxes_all``
(``/root/reference/mmdet3d/core/bbox/structures/base_box3d.py:539-568``),
``get_prediction_single`` (``mmdet3d/models/dense_heads/centerpoint_head_gga.py:250-341``) and
the consistency loss (``pgd_head.py:744-748`` / ``centerpoint_head_gga.py:714-720``) followed by
``loss.backward()``.  The step talks to the C ABI directly (no autograd graph): the loss
kernel emits ``d loss / d boxes`` in the same launch, scaled by ``loss_weight / avg_factor``
(mmdet ``weight_reduce_loss``).

``GeometryStep`` owns the output buffers of one batch shape so that a step is allocation
free and can be captured in a CUDA graph (``capture()`` / ``replay()``).
``GeometryStep.run_host`` is the same step for HOST buffers (pinned numpy / torch CPU
tensors): H2D copies, the kernels, and D2H of masks, loss and gradients.
"""
import ctypes

import torch

from . import _lib
from .losses import KINDS
from .project import MODES


class GeometryStep:

    def __init__(self, num_frames, num_points, num_boxes, device, pts_stride=4, kind='giou',
                 mode='lidar_direct', loss_weight=1.0, eps=1e-6, depth_clamp=0.1):
        self.F, self.N, self.M = int(num_frames), int(num_points), int(num_boxes)
        self.device = torch.device(device)
        self.pts_stride = int(pts_stride)
        self.kind = KINDS[kind]
        self.mode = MODES[mode]
        self.loss_weight, self.eps, self.depth_clamp = float(loss_weight), float(eps), float(depth_clamp)
        L = _lib.load()
        self.L = L
        self.W = L.gga_pib_row_words(self.M)
        dev = self.device
        n = self.F * self.M
        self.bits = torch.empty((self.F, self.N, self.W), dtype=torch.int32, device=dev)
        self.box2d = torch.empty((n, 4), dtype=torch.float32, device=dev)
        self.loss = torch.empty((n, 4 if self.kind == _lib.LOSS_L1 else 1), dtype=torch.float32, device=dev)
        self.loss_sum = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.loss_accum = None   # optional [1] running total shared by several steps (set by the caller)
        self.grad_boxes = torch.empty((n, 7), dtype=torch.float32, device=dev)
        # scratch of the loss reduction: owned by this step, so steps on parallel streams / graph
        # branches never share it (include/gga_b200.h: zeroed once, the kernel leaves it zeroed)
        self.scratch = torch.zeros((int(L.gga_loss_scratch_bytes()),), dtype=torch.uint8, device=dev)
        self.side = torch.cuda.Stream(device=dev)   # the box kernel runs beside the membership kernels
        self.graph = None
        self._host = None

    # ------------------------------------------------------------------ device-resident step
    def run(self, points, boxes, lidar2img, target, weight=None, avg_factor=None, after_loss=None):
        """points [F,N,pts_stride], boxes [F,M,7], lidar2img [F,M,4,4] (one calib per object,
        the GGA_lidar2img layout) or [4,4], target [F,M,4], weight [F,M] — contiguous fp32 CUDA
        tensors.  Enqueues the step on the current stream; returns nothing (results are in
        ``self.bits / box2d / loss / loss_sum / grad_boxes``; ``loss_sum`` is the weighted SUM,
        gradients are already scaled by ``loss_weight / avg_factor``).  ``after_loss`` (optional
        callable) runs on the side stream right after the loss kernel — the place for the scalar
        all-reduce of the loss sum (``gga_b200.dist.reduce_scalars``), which then overlaps the
        membership kernels and is captured with them in the CUDA graph."""
        L = self.L
        with torch.cuda.device(self.device):   # the C ABI works on the CURRENT device
            cur = torch.cuda.current_stream(self.device)
            st = cur.cuda_stream
            fork = torch.cuda.Event()
            fork.record(cur)
            self.side.wait_event(fork)
            _lib.check(L.gga_points_in_boxes_bits(points.data_ptr(), self.pts_stride, boxes.data_ptr(),
                                                  self.bits.data_ptr(), self.F, self.N, self.M, st),
                       'points_in_boxes_bits')
            a = self._box_args(boxes, lidar2img, target, weight, avg_factor)
            _lib.check(L.gga_box_project_loss(a, self.side.cuda_stream), 'box_project_loss')
            if after_loss is not None:
                with torch.cuda.stream(self.side):
                    after_loss(self)
            join = torch.cuda.Event()
            join.record(self.side)
            cur.wait_event(join)

    def capture(self, *args, **kw):
        """Warm up, then capture ``run(*args)`` into a CUDA graph (inputs are baked in by address)."""
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.run(*args, **kw)   # lazy initialisation must not happen under capture
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run(*args, **kw)
            self.graph = g
        return g

    def replay(self):
        self.graph.replay()

    # ------------------------------------------------------------------ host-buffer step
    def _box_args(self, boxes, lidar2img, target, weight, avg_factor):
        n = self.F * self.M
        a = _lib.BoxLossArgs()
        a.boxes = boxes.data_ptr()
        a.proj = lidar2img.data_ptr()
        a.proj_stride = 16 if lidar2img.dim() > 2 else 0
        a.target = target.data_ptr()
        if weight is not None:
            a.weight, a.weight_cols = weight.data_ptr(), 1
        a.n, a.mode, a.loss_kind = n, self.mode, self.kind
        a.depth_clamp, a.eps = self.depth_clamp, self.eps
        a.grad_scale = self.loss_weight / float(avg_factor if avg_factor is not None else max(n, 1))
        a.box2d, a.loss, a.loss_sum = self.box2d.data_ptr(), self.loss.data_ptr(), self.loss_sum.data_ptr()
        a.grad_boxes = self.grad_boxes.data_ptr()
        a.scratch, a.scratch_bytes = self.scratch.data_ptr(), self.scratch.numel()
        if self.loss_accum is not None:
            a.loss_accum = self.loss_accum.data_ptr()
        return a

    def _host_ctx(self, n_streams):
        if self._host is None or self._host['n_streams'] != n_streams:
            self.close()
            ctx = _lib.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self.L.gga_step_create(self.F, self.N, self.M, self.pts_stride, int(n_streams),
                                                  ctypes.byref(ctx)), 'step_create')
            self._host = dict(
                ctx=ctx, n_streams=n_streams, in_flight=None,
                h_bits=torch.empty(self.bits.shape, dtype=torch.int32).pin_memory(),
                h_grad=torch.empty(self.grad_boxes.shape, dtype=torch.float32).pin_memory(),
                h_loss=torch.empty((1,), dtype=torch.float32).pin_memory())
        return self._host

    def run_host(self, points, boxes, lidar2img, target, weight, avg_factor=None, n_streams=3, masks_to_host=True):
        """Same step with HOST inputs (page-locked torch CPU tensors) and HOST results: returns
        (bits_host int32 [F,N,W], loss_sum float, grad_boxes_host [F*M,7]).  Synchronous.
        ``masks_to_host='hits'`` returns them as int32 [n_hits, 2] = (frame * N + point, box) pairs
        (order unspecified) instead of dense rows: ~25x fewer bytes over PCIe at KITTI densities.
        ``masks_to_host=False`` leaves the masks on the device (the training use: the head consumes
        them there) and returns a CUDA int32 tensor view of the library's buffer instead — valid
        until the next ``run_host`` / ``close``.

        One C call (``gga_step_run_host``): the PCIe link is the bound (46 MB per step at the
        training shape), so the library pipelines the step frame by frame over `n_streams`
        streams — the H2D copy of frame f+1, the membership kernels of frame f and the D2H copy
        of the masks of frame f-1 overlap (full-duplex link)."""
        self.submit_host(points, boxes, lidar2img, target, weight, avg_factor, n_streams, masks_to_host)
        return self.wait_host()

    def submit_host(self, points, boxes, lidar2img, target, weight, avg_factor=None, n_streams=3, masks_to_host=True):
        """Asynchronous half of ``run_host``: enqueues the step (copies included) and returns.  The
        input tensors must stay alive and unmodified until ``wait_host()``, which returns what
        ``run_host`` returns.  One step in flight per ``GeometryStep``; two of them used alternately
        (``a.submit_host(batch k+1)`` before ``b.wait_host()`` of batch k) overlap the H2D copies
        of one batch with the D2H copies of the previous one — the PCIe link is full duplex."""
        L = self.L
        h = self._host_ctx(n_streams)
        for t in (points, boxes, lidar2img, target, weight):
            assert t is None or (not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()), \
                'run_host takes contiguous fp32 CPU tensors'
        cap = 0
        if masks_to_host == 'hits':
            if 'h_hits' not in h:
                n_cap = min(max(4096, 8 * self.F * self.N), 1 << 26)   # 8 hits per point on average (dense indoor scenes ~5.5)
                h['h_hits'] = torch.empty((n_cap, 2), dtype=torch.int32).pin_memory()
                h['h_nhits'] = torch.zeros((1,), dtype=torch.int32).pin_memory()
            cap = h['h_hits'].shape[0]
        n = self.F * self.M
        with torch.cuda.device(self.device):
            _lib.check(L.gga_step_submit_host(
                h['ctx'], points.data_ptr(), boxes.data_ptr(), lidar2img.data_ptr(), target.data_ptr(),
                None if weight is None else weight.data_ptr(), self.mode, self.kind, self.loss_weight,
                float(avg_factor if avg_factor is not None else max(n, 1)), self.eps, self.depth_clamp,
                h['h_bits'].data_ptr() if masks_to_host is True else None, cap, h['h_loss'].data_ptr(),
                h['h_grad'].data_ptr()), 'step_submit_host')
        h['in_flight'] = (masks_to_host, (points, boxes, lidar2img, target, weight))   # keeps the inputs alive

    def wait_host(self):
        h = self._host
        assert h is not None and h.get('in_flight') is not None, 'no step in flight'
        masks_to_host, _ = h['in_flight']
        h['in_flight'] = None
        hits = masks_to_host == 'hits'
        with torch.cuda.device(self.device):
            _lib.check(self.L.gga_step_wait_host(h['ctx'], h['h_hits'].data_ptr() if hits else None,
                                                 h['h_nhits'].data_ptr() if hits else None), 'step_wait_host')
        if hits:
            self.last_hits = int(h['h_nhits'][0])
            return h['h_hits'][:self.last_hits], float(h['h_loss'][0]), h['h_grad']
        if not masks_to_host:
            return self._device_bits(), float(h['h_loss'][0]), h['h_grad']
        return h['h_bits'], float(h['h_loss'][0]), h['h_grad']

    def _device_bits(self):
        h = self._host
        if 'd_bits' not in h:
            ptr = _lib.c_void_p()
            _lib.check(self.L.gga_step_device_bits(h['ctx'], ctypes.byref(ptr)), 'step_device_bits')

            class _Buf:   # zero-copy view of the library-owned device buffer
                __cuda_array_interface__ = {'shape': tuple(self.bits.shape), 'typestr': '<i4', 'data': (ptr.value, False),
                                            'version': 2}
            h['d_bits_owner'] = _Buf()
            h['d_bits'] = torch.as_tensor(h['d_bits_owner'], device=self.device)
        return h['d_bits']

    def close(self):
        """Releases the device context of the host-buffer path (idempotent)."""
        if self._host is not None:
            self.L.gga_step_destroy(self._host['ctx'])
            self._host = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def host_bytes(self, points, boxes, lidar2img, target, weight, masks_to_host=True):
        """(h2d, d2h) bytes moved by one ``run_host``."""
        h2d = sum(t.numel() * t.element_size() for t in (points, boxes, lidar2img, target, weight))
        if masks_to_host == 'hits':   # (row, box) pairs of the last run + their count
            mask_bytes = 8 * getattr(self, 'last_hits', 0) + 4
        else:
            mask_bytes = self.bits.numel() * 4 if masks_to_host else 0
        d2h = mask_bytes + self.grad_boxes.numel() * 4 + 4
        return h2d, d2h
