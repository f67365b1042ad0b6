#!/usr/bin/env python
"""bench.py — GGA geometry hot path: membership + projection + IoU/GIoU loss fwd+bwd, frames/s.

    python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3|c4|c5]
    python bench.py --impl reference --gpus N --steps K --warmup W [--workload ...]
    (N > 1: launched by torchrun, one rank per GPU)

Workloads (BASELINE.json `configs`; default c2 = the configuration the metric is quoted on):
  c1  KITTI single frame: 1 frame x 120 000 points x 64 proposals (the reference's CPU-runnable case)
  c2  GGA KITTI training shape: 8 frames per GPU per step, 120 000 points, 256 proposals per frame
  c3  FCAF3D + GGA SUN-RGBD shape: 8 frames per GPU per step, 50 000 points, 512 proposals, depth2img
  c4  pseudo-label matching pass: 464 frames per GPU x 512 detections vs 8 2D boxes, variant-B
      projection + pairwise IoU + argmax, match results all-gathered across ranks
  c5  roofline stress: 2 000 000 points x 1024 boxes; --partition replicate (one frame per rank, weak
      scaling) or split (the frame's points split across ranks, strong scaling)
c1/c2/c3/c5: one "step" = membership masks + variant-A projection + GIoU consistency loss, forward and
backward to the box parameters, over one batch of frames.  Synthetic data (gga_b200/synth.py,
SURVEY.md §8d).  Frames shard across ranks with no data-path collective; the only exchanges are the
ones the reference does: a scalar all-reduce of the accumulated loss sum at the log interval
(reduce_mean / _parse_losses + TextLoggerHook interval=50), issued asynchronously, and for c4 the
all-gather of the per-frame match results.

Prints ONE JSON line (rank 0).  Keys: see DESIGN.md §6.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = 'frames/s'
L2_BYTES = 126 * 1024 * 1024
WORKLOADS = {'c1': 1, 'c2': 2, 'c3': 3, 'c4': 4, 'c5': 5}
METRICS = {1: 'geometry_loss_fwd_bwd_frames_per_s', 2: 'geometry_loss_fwd_bwd_frames_per_s',
           3: 'geometry_loss_fwd_bwd_frames_per_s', 4: 'pseudo_label_matching_frames_per_s',
           5: 'geometry_loss_fwd_bwd_frames_per_s'}
WORKLOAD_NAMES = {1: 'kitti_single_frame (BASELINE.json configs[0])', 2: 'gga_kitti_train (BASELINE.json configs[1])',
                  3: 'fcaf3d_sunrgbd (BASELINE.json configs[2])', 4: 'pseudo_label_matching (BASELINE.json configs[3])',
                  5: 'roofline_stress (BASELINE.json configs[4])'}
GRAPH_CHUNK = 512   # steps per captured graph: --steps K is served by K // chunk replays + one graph of K % chunk steps


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def step_bytes(F, N, M, W):
    """Algorithmic bytes of one step (SURVEY.md §8d / BASELINE.md §3)."""
    member = F * (16 * N + 28 * M + 4 * N * W)
    boxes = F * M * ((28 + 64 + 16 + 16) + 32)
    return member, member + boxes


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                nv, 'nvmlDeviceGetCurrentClocksEventReasons') else int(
                nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
                     0x4: 'sw_power_cap', 0x80: 'hw_power_brake_slowdown'}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(self.period)

    def stop(self):
        self._stop.set()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': 0}
        return {'sm_mhz': int(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def bind_near_gpu(index):
    """Pins this process to the CPUs NVML names as closest to its GPU (same NUMA node / PCIe root), BEFORE
    any page-locked host buffer is allocated: with 8 ranks on one box the host-buffer (`e2e`) path is
    bound by host memory and PCIe root traffic, and first-touch then places the pinned pages next to
    the GPU.  Best effort: returns a short description for the JSON line, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        nv.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f'{len(cpus)} cpus [{cpus[0]}..{cpus[-1]}]'
    except Exception:
        return None


def physical_gpu_index(local):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local])
        except Exception:
            return local
    return local


def workload_config(cfg, c, n_gpus, partition='replicate'):
    out = {'workload': WORKLOAD_NAMES[cfg], 'frames_per_gpu_per_step': c['frames_per_gpu'],
           'boxes_per_frame': c['M'], 'parallelism': f'frames sharded x{n_gpus}'}
    if cfg == 4:
        out.update({'gt_boxes_per_frame': c['G'], 'pass': 'variant-B projection + pairwise 2D IoU + argmax; matches all-gathered once per 8 passes (one KITTI split)',
                    'global_frames_per_step': c['frames_per_gpu'] * n_gpus})
        return out
    out.update({'points_per_frame': c['N'], 'loss': 'giou(lidar_direct projection), fwd+bwd'})
    if cfg == 5 and partition == 'split':
        out['parallelism'] = f'the points of the frame split x{n_gpus} (boxes replicated)'
        out['global_frames_per_step'] = 1
    else:
        out['global_frames_per_step'] = c['frames_per_gpu'] * n_gpus
    return out


# ----------------------------------------------------------------------------- CPU reference path
def cpu_frame_fn(nthreads):
    """Returns f(frame_dict, n_points) running the reference's CPU path for one frame:
    points_in_boxes_cpu (oracle/pib_oracle.c, the literal mmcv loop) + torch-CPU corners ->
    lidar2img -> min/max -> GIoU loss -> backward (oracle/geometry.py, oracle/losses.py)."""
    import torch
    from oracle import geometry as og
    from oracle import losses as ol
    from oracle import membership as om
    om.build()
    torch.set_num_threads(max(1, nthreads))

    def run(f, n_points):
        pts = f['points'][:n_points]
        mask = om.points_in_boxes_all_np(pts, f['boxes'], nthreads=nthreads)
        M = f['boxes'].shape[0]
        b = torch.from_numpy(f['boxes']).clone().requires_grad_(True)
        l2i = torch.from_numpy(f['lidar2img'])[None].expand(M, 4, 4)
        box2d = og.project_lidar_direct(b, l2i)
        loss = ol.giou_loss_module(box2d, torch.from_numpy(f['target']), torch.from_numpy(f['weight']),
                                   avg_factor=float(M))
        loss.backward()
        return int(mask.sum()), float(loss.detach())
    return run


def cpu_match_fn(nthreads):
    """c4 on the CPU: the reference's per-frame conversion (oracle/geometry.project_kitti_cam =
    convert_valid_bboxes, kitti_dataset_GGA_match.py:685-765) + image_box_overlap + argmax
    (oracle/losses.py, eval.py:85-114 / utils_pseudo_labels_gga.py:45-60)."""
    import torch
    from oracle import geometry as og
    from oracle import losses as ol
    from gga_b200 import synth
    torch.set_num_threads(max(1, nthreads))
    rect, trv, p2 = (torch.from_numpy(x) for x in (synth.KITTI_RECT, synth.KITTI_TRV2C, synth.KITTI_P2))

    def run(f, _n=None):
        out = og.project_kitti_cam(torch.from_numpy(f['boxes']), rect, trv, p2, synth.KITTI_IMG_HW,
                                   synth.KITTI_MATCH_RANGE)
        box2d = out[0] if isinstance(out, (tuple, list)) else out
        ov = ol.image_box_overlap(np.asarray(box2d, dtype=np.float64), f['gt2d'].astype(np.float64))
        return int(ov.argmax(-1).sum()), 0.0
    return run


def numba_rbbox_leg(f, budget_s=4.0):
    """Second CPU leg: the reference's own numba points_in_rbbox
    (/root/reference/mmdet3d/core/bbox/box_np_ops.py:353-376) on the same frame, when the reference
    tree is reachable (it is not on the GPU box)."""
    try:
        from oracle import ref_loader
        ref = ref_loader.load_reference(None, None)
        fn = ref.box_np_ops.points_in_rbbox
    except Exception as e:  # noqa: BLE001
        return {'available': False, 'why': f'reference tree absent on this box ({type(e).__name__})'}
    pts, boxes = f['points'][:, :3], f['boxes']
    fn(pts[:1000], boxes)   # numba compile
    n, t0 = 0, time.perf_counter()
    while True:
        fn(pts, boxes)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 20:
            break
    return {'available': True, 'frames_per_s_membership_only': round(n / dt, 3), 'cores': 1,
            'sample': f'{n} frames of {pts.shape[0]} pts x {boxes.shape[0]} boxes, numba serial loop'}


def time_cpu_baseline(synth, cfg, budget_s=12.0):
    nthreads = os.cpu_count() or 1
    c = synth.CONFIGS[cfg]
    if cfg == 4:
        run = cpu_match_fn(nthreads)
        frames = [synth.make_frame(4, 900000 + i) for i in range(16)]
        run(frames[0])
        n, t0 = 0, time.perf_counter()
        while True:
            run(frames[n % 16])
            n += 1
            dt = time.perf_counter() - t0
            if dt >= budget_s or n >= 2000:
                break
        return {'value': round(n / dt, 3), 'unit': UNIT, 'cores': nthreads, 'kind': 'port',
                'sample': f'{n} frames of {c["M"]} detections x {c["G"]} 2D boxes (torch-CPU variant-B projection + '
                          f'numpy image_box_overlap + argmax), {dt:.1f} s'}
    run = cpu_frame_fn(nthreads)
    f = synth.make_frame(cfg, 900000)
    n_pts = c['N'] if cfg != 5 else 200000     # the stress frame is sampled: 10 % of its points
    run(f, min(n_pts, 20000))  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        run(f, n_pts)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 50:
            break
    out = {'value': round(n * (n_pts / c['N']) / dt, 3), 'unit': UNIT, 'cores': nthreads, 'kind': 'port',
           'sample': f'{n} x {n_pts} of {c["N"]} pts x {c["M"]} boxes (oracle/pib_oracle.c with '
                     f'{nthreads} OpenMP threads + torch-CPU projection/GIoU fwd+bwd), {dt:.1f} s'}
    if cfg in (1, 2):
        out['numba_points_in_rbbox'] = numba_rbbox_leg(f)
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from gga_b200 import synth
    cfg = WORKLOADS[args.workload]
    nthreads = os.cpu_count() or 1
    c = synth.CONFIGS[cfg]
    N, M = c['N'], c['M']
    frames = [synth.make_frame(cfg, 900000 + i) for i in range(2)]
    if cfg == 4:
        run = cpu_match_fn(nthreads)
        per_step = 1
        run(frames[0])
        t0 = time.perf_counter()
        run(frames[1])
        t1 = time.perf_counter() - t0
        per_step = int(max(1, min(c['frames_per_gpu'], 150.0 / max((args.steps + args.warmup) * t1, 1e-9))))
        fr = [synth.make_frame(4, 900000 + i) for i in range(min(per_step, 32))]
        for _ in range(args.warmup):
            for k in range(per_step):
                run(fr[k % len(fr)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for k in range(per_step):
                run(fr[k % len(fr)])
        dt = time.perf_counter() - t0
        fps = args.steps * per_step / dt
        sample = f'each step = {per_step} of {c["frames_per_gpu"]} frames x {M} detections x {c["G"]} 2D boxes'
    else:
        run = cpu_frame_fn(nthreads)
        run(frames[0], min(N, 20000))
        probe = min(N, 120000)
        t0 = time.perf_counter()
        run(frames[1], probe)
        t1 = (time.perf_counter() - t0) * (N / probe)
        total = args.steps + args.warmup
        frac = min(1.0, 150.0 / max(total * t1, 1e-9))
        n_pts = int(max(2000, min(N, round(frac * N))))
        for i in range(args.warmup):
            run(frames[i % 2], n_pts)
        t0 = time.perf_counter()
        for i in range(args.steps):
            run(frames[i % 2], n_pts)
        dt = time.perf_counter() - t0
        fps = args.steps * (n_pts / N) / dt
        sample = (f'each step = first {n_pts} of {N} points x {M} boxes of one frame (membership) + all {M} '
                  f'boxes projection/GIoU fwd+bwd; frames counted as {n_pts}/{N} per step')
    out = {
        'impl': 'reference', 'metric': METRICS[cfg], 'value': round(fps, 3), 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(dt / max(args.steps, 1) * 1e3, 4),
        'higher_is_better': True, 'scaling': 'strong' if (cfg == 5 and args.partition == 'split') else 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(cfg, c, args.gpus, args.partition),
        'cpu_baseline': {'value': round(fps, 3), 'unit': UNIT, 'cores': nthreads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': round(fps, 3), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class Ctx:
    pass


def setup_dist():
    import torch
    import torch.distributed as dist
    x = Ctx()
    x.rank = int(os.environ.get('RANK', '0'))
    x.world = int(os.environ.get('WORLD_SIZE', '1'))
    x.local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    torch.cuda.set_device(x.local)
    x.dev = torch.device('cuda', x.local)
    # (only with several ranks: the single-rank run also times the CPU baseline on ALL host cores)
    x.numa = bind_near_gpu(physical_gpu_index(x.local)) if x.world > 1 else None
    if x.world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=x.dev, timeout=datetime.timedelta(seconds=180))
    x.dist = dist

    def barrier():
        if x.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    x.barrier = barrier

    def max_over_ranks(ms):
        if x.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=x.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms
    x.max_over_ranks = max_over_ranks
    return x


def ramp_reps(x, fn, seconds=0.25, sync=None):
    """How often to repeat fn() so that ~`seconds` of this work run before the timed region (SM clock
    ramp).  fn may hold collectives, so every rank must repeat it the SAME number of times: the count
    is derived from one probe run and agreed on through a max over ranks."""
    import math
    if sync is None:
        import torch
        sync = torch.cuda.synchronize
    sync()
    t0 = time.perf_counter()
    fn()
    sync()
    t1 = x.max_over_ranks(time.perf_counter() - t0)
    return int(min(5000, max(1, math.ceil(seconds / max(t1, 1e-6)))))


def timed(x, fn, sampler=None):
    """barrier + sync, CUDA events around fn() on the current stream, barrier + sync; max over ranks (ms)."""
    import torch
    x.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    if sampler is not None:
        sampler.sample()   # the GPU is still busy with the timed work here
    x.barrier()
    return x.max_over_ranks(e0.elapsed_time(e1))


def run_steps_workload(args, x, cfg):
    import torch
    import gga_b200 as G
    from gga_b200 import synth
    from gga_b200.step import GeometryStep
    dist, dev, world, rank = x.dist, x.dev, x.world, x.rank
    c = synth.CONFIGS[cfg]
    F, N, M = c['frames_per_gpu'], c['N'], c['M']
    split = cfg == 5 and args.partition == 'split' and world > 1
    N_full = N
    if split:   # the single frame's points are split across the ranks (boxes replicated)
        from gga_b200 import dist as gdist
        lo, hi = gdist.shard_range(N, rank, world)
        N = hi - lo
    W = G.row_words(M)
    member_bytes, all_bytes = step_bytes(F, N, M, W)
    n_sets = max(3, -(-2 * L2_BYTES // all_bytes) + 1)   # rotating working set > 2x L2
    lanes = max(1, min(args.lanes, n_sets))
    n_sets = -(-n_sets // lanes) * lanes                 # every buffer set belongs to one lane
    host = []
    for k in range(2):
        hb = synth.make_batch(cfg, 100000 * (0 if split else rank) + 16 * k, F)
        if split:
            hb['points'] = np.ascontiguousarray(hb['points'][:, lo:hi])
        host.append(hb)
    # loss scalars: every step adds its weighted loss sum to a device accumulator (inside the loss
    # kernel); the accumulator is all-reduced across ranks at the log interval, asynchronously on a
    # side stream (the reference: mmdet reduce_mean / _parse_losses feeding a TextLoggerHook with
    # interval=50, configs/gga/gga_kitti_config.py:251-254).  No collective sits on the data path.
    acc = torch.zeros((4,), dtype=torch.float32, device=dev)
    snaps = torch.zeros((64, 4), dtype=torch.float32, device=dev)
    comm = torch.cuda.Stream() if world > 1 else None
    works = []
    if world > 1:   # create the NCCL communicator up front
        dist.all_reduce(snaps[0])
        torch.cuda.synchronize()

    def reduce_scalars_async():
        ev = torch.cuda.Event()
        ev.record()
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            j = len(works) % 64
            snaps[j].copy_(acc)
            works.append(dist.all_reduce(snaps[j], async_op=True))

    sets, steps = [], []
    names = ('points', 'boxes', 'lidar2img', 'target', 'weight')
    for k in range(n_sets):
        hb = host[k % 2]
        t = {name: torch.from_numpy(np.ascontiguousarray(hb[name])).to(dev) for name in names}
        sets.append(t)
        s = GeometryStep(F, N, M, dev, kind='giou', mode='lidar_direct')
        s.loss_accum = acc[0:1]
        s.run(t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(F * M))   # warm-up (lazy init)
        steps.append(s)
    torch.cuda.synchronize()
    lane_streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)] if lanes > 1 else []

    def capture_steps(n, first=0, n_lanes=lanes):
        """ONE graph of exactly n steps (n full passes, each over its own buffer set); with lanes the
        steps are issued round-robin on parallel branches (independent batches in flight together)."""
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream()
            if n_lanes > 1:
                fork = torch.cuda.Event()
                fork.record(cur)
                for ls in lane_streams[:n_lanes]:
                    ls.wait_event(fork)
            for i in range(n):
                k = (first + i) % n_sets
                t = sets[k]
                if n_lanes > 1:
                    with torch.cuda.stream(lane_streams[k % n_lanes]):
                        steps[k].run(t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(F * M))
                else:
                    steps[k].run(t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(F * M))
            for ls in lane_streams[:n_lanes] if n_lanes > 1 else []:
                cur.wait_stream(ls)
        torch.cuda.synchronize()
        return g

    def plan(K):
        """graphs that together run exactly K steps: K // chunk replays of a chunk graph + one remainder graph"""
        chunk = min(K, GRAPH_CHUNK)
        if K > GRAPH_CHUNK:      # several replays: whole rotations keep the sets evenly used
            chunk -= chunk % n_sets
        chunk = max(chunk, 1)    # (K <= GRAPH_CHUNK: ONE graph holds all K steps — one launch in the timed region)
        full, rem = divmod(K, chunk)
        g_chunk = capture_steps(chunk)
        g_rem = capture_steps(rem, first=(full * chunk) % n_sets) if rem else None
        return full, g_chunk, g_rem

    log_every = max(1, 50 // max(1, min(args.steps, GRAPH_CHUNK)))   # ~ every 50 steps

    def make_runner(K):
        full, g_chunk, g_rem = plan(K)

        def run():
            for i in range(full):
                g_chunk.replay()
                if world > 1 and (i + 1) % log_every == 0:
                    reduce_scalars_async()
            if g_rem is not None:
                g_rem.replay()
            if world > 1:
                reduce_scalars_async()
        return run

    warm = make_runner(max(args.warmup, 3))
    main_run = make_runner(args.steps)
    # warm-up: at least W steps, and at least ~0.25 s of this same work so that the clocks have ramped
    # (the driver's default run times only 20 steps = 0.3 ms of GPU work after a long host set-up)
    warm()
    for _ in range(ramp_reps(x, main_run)):
        main_run()
        torch.cuda.synchronize()
    for wk in works:
        wk.wait()
    del works[:]
    sampler = ClockSampler(physical_gpu_index(x.local))
    sampler.start()
    x.barrier()
    sampler.sample()
    ms = timed(x, main_run, sampler)
    if comm is not None:
        torch.cuda.current_stream().wait_stream(comm)
    for wk in works:
        wk.wait()
    frames_step_global = 1 if split else world * F
    value = frames_step_global * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps

    # the same K steps strictly one after the other (one lane): the sequential-step number
    seq = None
    if lanes > 1:
        Ks = min(args.steps, 200)
        full, g_chunk, g_rem = None, None, None
        g_seq = capture_steps(Ks, n_lanes=1)
        g_seq.replay()
        sms = timed(x, g_seq.replay)
        seq = {'value': round(frames_step_global * Ks / (sms * 1e-3), 1), 'unit': UNIT, 'steps': Ks,
               'ms_per_step': round(sms / Ks, 5), 'frames_in_flight_per_gpu': F}
        del g_seq
    sampler.stop()

    # dominant kernel alone (membership), same rotation of buffers, CUDA events on the launch stream
    L = G._lib.load()

    def member(i):
        t, s = sets[i % n_sets], steps[i % n_sets]
        st = torch.cuda.current_stream().cuda_stream
        rc = L.gga_points_in_boxes_bits(t['points'].data_ptr(), 4, t['boxes'].data_ptr(), s.bits.data_ptr(), F, N, M, st)
        assert rc == 0
    for i in range(n_sets):
        member(i)
    torch.cuda.synchronize()
    mg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(mg):       # one rotation of membership launches, replayed by the driver
        for i in range(n_sets):
            member(i)
    mg.replay()
    torch.cuda.synchronize()
    kreps = max(8, min(args.steps, 500) // n_sets)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(kreps):
        mg.replay()
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / (kreps * n_sets)
    peak, peak_src = peaks()
    achieved = member_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        if tj.get('workload', 'c2') == args.workload:
            traffic = int(tj['membership_dram_bytes_per_launch'])
    except Exception:
        pass

    # end to end through the public API with HOST buffers (pinned), copies inside the timed region
    # Two batches alternate (own pinned inputs, own GeometryStep = own device context and pinned
    # result buffers): `pipelined` submits batch k+1 before it waits for batch k, so the H2D copies of
    # one batch overlap the D2H copies of the previous one (what a prefetching data loader does);
    # `synchronous` is one blocking run_host call per step.
    depth = max(2, args.e2e_depth)
    hins = [{name: torch.from_numpy(np.ascontiguousarray(host[k % len(host)][name])).pin_memory() for name in names}
            for k in range(2)]
    ess = [GeometryStep(F, N, M, dev, kind='giou', mode='lidar_direct') for _ in range(depth)]
    eargs = [(h['points'], h['boxes'], h['lidar2img'], h['target'], h['weight'], float(F * M)) for h in hins]
    ereps = max(3 * depth, min(args.steps, 40))
    es = ess[0]

    def e2e_leg(pipelined=True, **kw):
        # pipelined steps overlap each other, so each is ONE copy each way + one launch (n_streams = 1);
        # a blocking step overlaps its own frames over a few streams instead
        kw = dict(kw, n_streams=args.e2e_pipe_streams if pipelined else args.e2e_streams)
        for k in range(2 * depth):
            ess[k % depth].run_host(*eargs[k % 2], **kw)

        def loop_sync():
            for k in range(ereps):
                ess[k % depth].run_host(*eargs[k % 2], **kw)

        def loop_pipe():
            for k in range(ereps):
                if k >= depth:
                    ess[k % depth].wait_host()          # batch k - depth
                ess[k % depth].submit_host(*eargs[k % 2], **kw)
            for k in range(ereps, ereps + depth):
                ess[k % depth].wait_host()
        ems = timed(x, loop_pipe if pipelined else loop_sync)
        return round(frames_step_global * ereps / (ems * 1e-3), 1), round(ems / ereps, 4)
    ev, ems_step = e2e_leg()
    evs, ems_sync = e2e_leg(pipelined=False)
    h2d, d2h = es.host_bytes(*eargs[0][:5])
    e2e = {'value': ev, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'steps': ereps,
           'ms_per_step': ems_step, 'pipeline_depth': depth,
           'returns': 'masks + loss + box gradients to host memory (the points_in_boxes_cpu-style contract)',
           'synchronous': {'value': evs, 'unit': UNIT, 'ms_per_step': ems_sync}}
    # same call with the masks left on the device (the training use: only loss and gradients go back)
    ev2, ems2 = e2e_leg(masks_to_host=False)
    ev2s, ems2s = e2e_leg(pipelined=False, masks_to_host=False)
    e2e['masks_on_device'] = {'value': ev2, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                              'd2h_bytes_per_step': int(es.host_bytes(*eargs[0][:5], masks_to_host=False)[1]),
                              'ms_per_step': ems2, 'synchronous': {'value': ev2s, 'ms_per_step': ems2s}}
    # ... and with the masks returned as a compact hit list (point, box) instead of dense bit rows
    try:
        ev3, ems3 = e2e_leg(masks_to_host='hits')
        ev3s, ems3s = e2e_leg(pipelined=False, masks_to_host='hits')
        e2e['masks_as_hit_list'] = {'value': ev3, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                                    'd2h_bytes_per_step': int(es.host_bytes(*eargs[0][:5], masks_to_host='hits')[1]),
                                    'ms_per_step': ems3, 'synchronous': {'value': ev3s, 'ms_per_step': ems3s}}
    except (TypeError, NotImplementedError, RuntimeError) as e:   # e.g. the list did not fit its capacity
        e2e['masks_as_hit_list'] = {'unavailable': str(e)[:160]}
    for s_ in ess:
        s_.close()

    # the same membership through the mmcv-layout entry point (int32 [F, N, M], what a drop-in
    # `points_in_boxes_all` caller gets), reported beside the step's own kernel (never part of `value`)
    mmcv_layout = None
    if F * N * M * 4 * 2 < 40e9:
        try:
            a_outs = [torch.empty((F, N, M), dtype=torch.int32, device=dev) for _ in range(2)]

            def member_all(i):
                t = sets[i % n_sets]
                rc = L.gga_points_in_boxes_all(t['points'].data_ptr(), 4, t['boxes'].data_ptr(), a_outs[i % 2].data_ptr(),
                                               F, N, M, torch.cuda.current_stream().cuda_stream)
                assert rc == 0
            for i in range(2):
                member_all(i)
            torch.cuda.synchronize()
            ag = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ag):
                for i in range(4):
                    member_all(i)
            ag.replay()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(5):
                ag.replay()
            a1.record()
            torch.cuda.synchronize()
            all_ms = a0.elapsed_time(a1) / 20
            all_alg = F * (16 * N + 28 * M + 4 * N * M)
            mmcv_layout = {'entry': 'gga_points_in_boxes_all (int32 [F, N, M])', 'kernel_ms': round(all_ms, 5),
                           'algorithmic_bytes_per_launch': all_alg, 'achieved': round(all_alg / (all_ms * 1e-3) / 1e9, 1),
                           'frac': round(all_alg / (all_ms * 1e-3) / 1e9 / peak, 4)}
            del a_outs, ag
        except torch.cuda.OutOfMemoryError:
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = time_cpu_baseline(synth, cfg)

    if rank == 0:
        conf = workload_config(cfg, c, world, args.partition)
        if split:
            conf['points_per_gpu'] = N
            conf['points_per_frame'] = N_full
        protocol = {'l2': f'rotating {n_sets} input/output sets ({n_sets * all_bytes / 1e6:.0f} MB > 2x 126 MB L2)',
                    'launch': (f'exactly --steps steps inside the timed region: CUDA graphs of up to {GRAPH_CHUNK} steps; '
                               f'{lanes} step lane(s) = parallel graph branches, each buffer set bound to one lane'),
                    'frames_in_flight_per_gpu': F * lanes,
                    'warmup': 'W steps, then the timed graphs replayed for 0.25 s so that the SM clock has ramped'}
        out = {
            'metric': METRICS[cfg], 'value': round(value, 1), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': round(ms_per_step, 5), 'higher_is_better': True,
            'scaling': 'strong' if split else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': conf, 'protocol': dict(protocol, cpu_binding=x.numa), 'lanes': lanes, 'sequential_step': seq,
            'roofline': {'bound': 'hbm', 'kernel': 'pib_sweep_kernel (membership, bit-packed rows; one launch per call)',
                         'achieved': round(achieved, 1), 'peak': peak, 'unit': 'GB/s', 'frac': round(achieved / peak, 4),
                         'traffic': traffic, 'peak_source': peak_src, 'kernel_ms': round(kernel_ms, 5),
                         'algorithmic_bytes_per_launch': member_bytes,
                         'step_frac_of_hbm_roofline': round(all_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                         'mmcv_layout': mmcv_layout,
                         # the brute-force definition of the work (14 FP32 instructions per
                         # point-box pair) over the same launch time
                         'fp32_bruteforce_equivalent': {'flops_per_launch': 14 * F * N * M,
                                                        'tflops': round(14 * F * N * M / (kernel_ms * 1e-3) / 1e12, 1),
                                                        'peak_tflops': 74.4}},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': 2 * args.steps,
            'clocks': sampler.summary(),
        }
        print(json.dumps(out), flush=True)


def run_match_workload(args, x):
    """c4: one step = one pass over this rank's 464 frames: variant-B projection of 512 detections per
    frame, block-diagonal pairwise IoU against the frame's 2D boxes, argmax; the match indices and
    best IoUs of all ranks are all-gathered (the reference runs the pass on rank 0 over all frames,
    tools/generate_pseudo_labels_gga.py:242 / utils_pseudo_labels_gga.py:17-84)."""
    import torch
    import gga_b200 as G
    from gga_b200 import matching, synth
    dist, dev, world, rank = x.dist, x.dev, x.world, x.rank
    c = synth.CONFIGS[4]
    F, M, Gt = c['frames_per_gpu'], c['M'], c['G']
    base = [synth.make_frame(4, 5000 * rank + i) for i in range(16)]
    boxes_h = np.concatenate([base[i % 16]['boxes'] for i in range(F)]).astype(np.float32)      # [F*M, 7]
    gts = [base[i % 16]['gt2d'].astype(np.float64) for i in range(F)]
    gt_h = np.concatenate(gts)
    gt_off_h = np.concatenate([[0], np.cumsum([len(g) for g in gts])]).astype(np.int32)
    dt_off_h = (np.arange(F + 1) * M).astype(np.int32)
    n_sets = 3
    sets = []
    rect, trv, p2 = (torch.from_numpy(a).to(dev) for a in (synth.KITTI_RECT, synth.KITTI_TRV2C, synth.KITTI_P2))
    fob = torch.arange(F, device=dev, dtype=torch.int32).repeat_interleave(M)
    rect, trv, p2 = (m[None].expand(F, 4, 4).contiguous() for m in (rect, trv, p2))      # one calib per frame
    img_hw = torch.tensor(synth.KITTI_IMG_HW, dtype=torch.float32, device=dev)[None].expand(F, 2).contiguous()
    pcd = torch.tensor(synth.KITTI_MATCH_RANGE, dtype=torch.float32, device=dev)
    for k in range(n_sets):
        sets.append(dict(boxes=torch.from_numpy(boxes_h).to(dev) + 0.001 * k, gt=torch.from_numpy(gt_h).to(dev),
                         gt_off=torch.from_numpy(gt_off_h).to(dev), dt_off=torch.from_numpy(dt_off_h).to(dev)))

    def one_pass(k):
        s = sets[k % n_sets]
        conv = matching.convert_valid_bboxes_batch(s['boxes'], fob, rect, trv, p2, img_hw, pcd)
        match, best = matching.match_dt_to_gt(conv['bbox'], s['dt_off'], s['gt'], s['gt_off'])
        return match, best

    for k in range(n_sets):
        one_pass(k)
    torch.cuda.synchronize()
    graphs, outs = [], []
    for k in range(n_sets):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs.append(one_pass(k))
        graphs.append(g)
    torch.cuda.synchronize()
    # The exchange step of the pass: the reference rewrites the annos of the WHOLE split from the matches
    # (utils_pseudo_labels_gga.py:62-84), so every rank needs all of them once per split, not per batch:
    # matches are collected on the device for EPOCH passes (8 x 464 = the 3712 frames of the KITTI train
    # split) and then all-gathered in one NCCL call on a side stream, overlapped with the next passes.
    # Only the match INDEX travels (int16: lossless, a frame has < 32768 2D boxes): the rewrite consumes
    # nothing else (`dt_match_gt = np.argmax(c_overlap, axis=-1)`, :62); the best IoU stays on the rank.
    EPOCH = 8
    comm2 = torch.cuda.Stream() if world > 1 else None     # the all-gathers
    if world > 1:
        assert Gt < 32768
        acc2 = [torch.zeros((EPOCH, F * M), dtype=torch.int16, device=dev) for _ in range(2)]
        gathered = torch.empty((world, EPOCH, F * M), dtype=torch.int16, device=dev)
        gather_done = [None, None]
        dist.all_gather_into_tensor(gathered.view(torch.uint8).view(-1), acc2[0].view(torch.uint8).view(-1))   # communicator set-up (bytes: NCCL has no int16)
        torch.cuda.synchronize()

    graphs2 = None
    if world > 1:
        # one graph per (epoch buffer, slot): the pass + the int32 -> int16 pack of its match indices
        # straight into the epoch buffer, so a pass costs the host ONE graph launch (8 ranks share the
        # box's host cores: per-pass Python work would otherwise bound the pass)
        graphs2 = [[None] * EPOCH for _ in range(2)]
        for b in range(2):
            for slot in range(EPOCH):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    match, _best = one_pass(slot)
                    acc2[b][slot].copy_(match)
                graphs2[b][slot] = g
        torch.cuda.synchronize()

    def gather_epoch(b, cur):
        ev = torch.cuda.Event()
        ev.record(cur)
        comm2.wait_event(ev)
        with torch.cuda.stream(comm2):
            dist.all_gather_into_tensor(gathered.view(torch.uint8).view(-1), acc2[b].view(torch.uint8).view(-1))
            gather_done[b] = torch.cuda.Event()
            gather_done[b].record(comm2)

    def run(K):
        cur = torch.cuda.current_stream()
        if world == 1:
            for i in range(K):
                graphs[i % n_sets].replay()
            return
        for i in range(K):
            b, slot = (i // EPOCH) % 2, i % EPOCH
            if slot == 0 and gather_done[b] is not None:
                cur.wait_event(gather_done[b])   # the gather that read this epoch buffer is done
            graphs2[b][slot].replay()
            if slot == EPOCH - 1 or i == K - 1:
                gather_epoch(b, cur)
        cur.wait_stream(comm2)
    run(max(args.warmup, 3))
    for _ in range(ramp_reps(x, lambda: run(args.steps))):
        run(args.steps)
        torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(x.local))
    sampler.start()
    x.barrier()
    sampler.sample()
    ms = timed(x, lambda: run(args.steps), sampler)
    sampler.stop()
    value = world * F * args.steps / (ms * 1e-3)
    # device-only time of one pass (graph replays back to back, no collective): the kernels' own time
    kms = timed(x, lambda: [graphs[i % n_sets].replay() for i in range(20)]) / 20
    n = F * M
    alg = n * (28 + 16 + 1) + F * (Gt * 32) + n * 8 + 3 * 64
    peak, peak_src = peaks()
    # end to end with host buffers: boxes up, match + best IoU down
    hb = torch.from_numpy(boxes_h).pin_memory()
    hm = torch.empty((n,), dtype=torch.int32).pin_memory()
    hi = torch.empty((n,), dtype=torch.float32).pin_memory()
    dbox = torch.empty((n, 7), dtype=torch.float32, device=dev)

    def host_pass():
        dbox.copy_(hb, non_blocking=True)
        conv = matching.convert_valid_bboxes_batch(dbox, fob, rect, trv, p2, img_hw, pcd)
        match, best = matching.match_dt_to_gt(conv['bbox'], sets[0]['dt_off'], sets[0]['gt'], sets[0]['gt_off'])
        hm.copy_(match, non_blocking=True)
        hi.copy_(best, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(3):
        host_pass()
    ereps = max(5, min(args.steps, 40))
    ems = timed(x, lambda: [host_pass() for _ in range(ereps)])
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = time_cpu_baseline(synth, 4)
    if rank == 0:
        conf = workload_config(4, c, world)
        protocol = {'launch': 'one CUDA graph per pass (projection + matching kernels); NCCL all_gather_into_tensor of the collected matches every 8 passes on a side stream',
                    'l2': 'three rotating input sets (11 MB per pass: L2 resident, the pass is latency bound)'}
        out = {
            'metric': METRICS[4], 'value': round(value, 1), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 5), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (IoU in f64 like the reference)', 'data': 'synthetic',
            'config': conf, 'protocol': protocol,
            'roofline': {'bound': 'hbm', 'kernel': 'box_loss_kernel (variant-B projection) + match kernel, one pass',
                         'achieved': round(alg / (kms * 1e-3) / 1e9, 1), 'peak': peak, 'unit': 'GB/s',
                         'frac': round(alg / (kms * 1e-3) / 1e9 / peak, 4), 'traffic': None, 'peak_source': peak_src,
                         'kernel_ms': round(kms, 5), 'algorithmic_bytes_per_launch': alg,
                         'note': 'latency bound: 11 MB per pass, a few launches of one thread per box / one warp per detection row'},
            'cpu_baseline': cpu,
            'e2e': {'value': round(world * F * ereps / (ems * 1e-3), 1), 'unit': UNIT, 'h2d_bytes_per_step': int(n * 28),
                    'd2h_bytes_per_step': int(n * 8), 'steps': ereps, 'ms_per_step': round(ems / ereps, 4)},
            'gpu_launches': 'see profiles (projection + matching kernels per pass)', 'clocks': sampler.summary(),
        }
        print(json.dumps(out), flush=True)


def run_ours(args):
    x = setup_dist()
    cfg = WORKLOADS[args.workload]
    if cfg == 4:
        run_match_workload(args, x)
    else:
        run_steps_workload(args, x, cfg)
    if x.world > 1:
        x.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--partition', default='replicate', choices=['replicate', 'split'], help='c5 across ranks')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-streams', type=int, default=3, help='streams of the frame pipeline inside a blocking host step')
    ap.add_argument('--e2e-depth', type=int, default=2, help='host-buffer steps kept in flight (alternating contexts)')
    ap.add_argument('--e2e-pipe-streams', type=int, default=1, help='streams inside a pipelined host step (1 = one copy each way)')
    ap.add_argument('--lanes', type=int, default=2, help='independent steps in flight on one GPU (parallel graph branches)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
