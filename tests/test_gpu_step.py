"""GPU parity tests of the training-shaped step (gga_b200/step.py): membership bits, loss and
gradients of one GeometryStep against the CPU oracle; CUDA-graph replay and the host-buffer
(pipelined H2D / kernels / D2H) variant must give identical results."""
import numpy as np
import pytest
import torch

import gga_b200 as G
from gga_b200 import synth
from gga_b200.step import GeometryStep
from oracle import geometry as og
from oracle import losses as ol
from oracle import membership as om
from parity import close64

pytestmark = pytest.mark.gpu
F, N, M = 3, 6000, 200


def _batch():
    bt = synth.make_batch(2, 40, F, N=N, M=M)
    return {k: np.ascontiguousarray(bt[k]) for k in ('points', 'boxes', 'lidar2img', 'target', 'weight')}


def _oracle(bt):
    masks = np.stack([om.points_in_boxes_all_np(bt['points'][f], bt['boxes'][f], 8) for f in range(F)])
    b = torch.from_numpy(bt['boxes']).reshape(-1, 7).clone().requires_grad_(True)
    box2d = og.project_lidar_direct(b, torch.from_numpy(bt['lidar2img']).reshape(-1, 4, 4))
    loss = ol.giou_loss_module(box2d, torch.from_numpy(bt['target']).reshape(-1, 4),
                               torch.from_numpy(bt['weight']).reshape(-1), avg_factor=float(F * M))
    loss.backward()
    b64 = torch.from_numpy(bt['boxes']).reshape(-1, 7).double().requires_grad_(True)
    loss64 = ol.giou_loss_module(og.project_lidar_direct(b64, torch.from_numpy(bt['lidar2img']).reshape(-1, 4, 4).double()),
                                 torch.from_numpy(bt['target']).reshape(-1, 4).double(),
                                 torch.from_numpy(bt['weight']).reshape(-1).double(), avg_factor=float(F * M))
    loss64.backward()
    return masks, float(loss), b.grad.numpy(), float(loss64), b64.grad.numpy()


def _check(step, bits, loss_sum, grad, ref):
    masks, rloss, rgrad, loss64, grad64 = ref
    got = G.unpack_bits(torch.as_tensor(bits).cuda(), M).cpu().numpy()
    assert np.array_equal(got, masks)                                  # bit-exact
    loss = float(loss_sum) / (F * M)
    assert abs(loss - loss64) <= 1e-5 * abs(loss64) + 1e-7              # 1e-5 relative (north_star), float64 yardstick
    assert close64(np.asarray(grad), grad64, rgrad, what='step grad')


def test_step_matches_oracle_eager_graph_and_host():
    bt = _batch()
    ref = _oracle(bt)
    dev = torch.device('cuda:0')
    t = {k: torch.from_numpy(v).to(dev) for k, v in bt.items()}
    step = GeometryStep(F, N, M, dev, kind='giou', mode='lidar_direct')
    args = (t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(F * M))
    step.run(*args)
    torch.cuda.synchronize()
    _check(step, step.bits.cpu(), step.loss_sum.item(), step.grad_boxes.cpu(), ref)
    # CUDA graph: capture once, replay twice on scrubbed outputs
    step.capture(*args)
    for _ in range(2):
        step.bits.fill_(-1); step.grad_boxes.zero_(); step.loss_sum.zero_()
        step.replay()
        torch.cuda.synchronize()
        _check(step, step.bits.cpu(), step.loss_sum.item(), step.grad_boxes.cpu(), ref)
    # host buffers in, host results out
    hin = {k: torch.from_numpy(v).pin_memory() for k, v in bt.items()}
    hs = GeometryStep(F, N, M, dev, kind='giou', mode='lidar_direct')
    for _ in range(2):
        bits, loss_sum, grad = hs.run_host(hin['points'], hin['boxes'], hin['lidar2img'], hin['target'],
                                           hin['weight'], float(F * M))
        assert not bits.is_cuda and not grad.is_cuda
        _check(hs, bits, loss_sum, grad, ref)
    h2d, d2h = hs.host_bytes(hin['points'], hin['boxes'], hin['lidar2img'], hin['target'], hin['weight'])
    assert h2d == sum(v.nbytes for v in bt.values()) and d2h == F * N * step.W * 4 + F * M * 7 * 4 + 4
    # training use: masks stay on the device, only loss and gradients come back
    dbits, loss_sum, grad = hs.run_host(hin['points'], hin['boxes'], hin['lidar2img'], hin['target'],
                                        hin['weight'], float(F * M), masks_to_host=False)
    assert dbits.is_cuda and dbits.dtype == torch.int32 and not grad.is_cuda
    _check(hs, dbits.cpu(), loss_sum, grad, ref)
    assert hs.host_bytes(hin['points'], hin['boxes'], hin['lidar2img'], hin['target'], hin['weight'],
                         masks_to_host=False)[1] == F * M * 7 * 4 + 4


def test_hit_list_equals_dense_masks():
    """The compact (row, box) pair list — from device rows and through the host-buffer step — names
    exactly the set bits of the dense rows (order unspecified)."""
    bt = _batch()
    dev = torch.device('cuda:0')
    t = {k: torch.from_numpy(v).to(dev) for k, v in bt.items()}
    bits = G.points_in_boxes_bits(t['points'], t['boxes'])
    dense = G.unpack_bits(bits, M).cpu().numpy().reshape(F * N, M)
    want = np.stack(np.nonzero(dense), 1)
    got = G.hit_list(bits, M).cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.int32
    assert np.array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], want)
    with pytest.raises(RuntimeError):
        G.hit_list(bits, M, capacity=max(1, len(want) // 2))      # overflow is reported, not silent
    hin = {k: torch.from_numpy(v).pin_memory() for k, v in bt.items()}
    hs = GeometryStep(F, N, M, dev, kind='giou', mode='lidar_direct')
    ref = _oracle(bt)
    for _ in range(2):
        hits, loss_sum, grad = hs.run_host(hin['points'], hin['boxes'], hin['lidar2img'], hin['target'], hin['weight'],
                                           float(F * M), masks_to_host='hits')
        h = hits.numpy()
        assert np.array_equal(h[np.lexsort((h[:, 1], h[:, 0]))], want)
        assert abs(loss_sum - ref[1] * F * M) <= 1e-5 * abs(ref[1] * F * M) + 1e-6
    assert hs.host_bytes(hin['points'], hin['boxes'], hin['lidar2img'], hin['target'], hin['weight'],
                         masks_to_host='hits')[1] == 8 * len(want) + 4 + F * M * 7 * 4 + 4


@pytest.mark.parametrize('n_streams', [1, 3])
def test_pipelined_host_steps_equal_synchronous_ones(n_streams):
    """submit_host / wait_host: two contexts used alternately on two different batches (batch k+1 is
    submitted before batch k is waited for) return what the blocking call returns, for dense rows,
    masks left on the device and the hit list; a second submit on a busy context is refused."""
    dev = torch.device('cuda:0')
    bts = []
    for seed in (40, 41):
        bt = synth.make_batch(2, seed, F, N=N, M=M)
        bts.append({k: np.ascontiguousarray(bt[k]) for k in ('points', 'boxes', 'lidar2img', 'target', 'weight')})
    refs = [_oracle(bt) for bt in bts]
    hins = [{k: torch.from_numpy(v).pin_memory() for k, v in bt.items()} for bt in bts]
    args = [(h['points'], h['boxes'], h['lidar2img'], h['target'], h['weight'], float(F * M)) for h in hins]
    ss = [GeometryStep(F, N, M, dev, kind='giou', mode='lidar_direct') for _ in range(2)]
    for mode in (True, False, 'hits'):
        ss[0].submit_host(*args[0], masks_to_host=mode, n_streams=n_streams)
        for k in range(1, 6):
            ss[k % 2].submit_host(*args[k % 2], masks_to_host=mode, n_streams=n_streams)
            j = (k - 1) % 2
            out, loss_sum, grad = ss[j].wait_host()
            if mode == 'hits':
                h = out.numpy()
                dense = np.zeros((F * N, M), dtype=np.int32)
                dense[h[:, 0], h[:, 1]] = 1
                assert len(h) == refs[j][0].sum() and np.array_equal(dense.reshape(F, N, M), refs[j][0])
                assert abs(loss_sum / (F * M) - refs[j][3]) <= 1e-5 * abs(refs[j][3]) + 1e-7
            else:
                _check(ss[j], out.cpu() if out.is_cuda else out, loss_sum, grad, refs[j])
        ss[5 % 2].wait_host()
    ss[0].submit_host(*args[0], n_streams=n_streams)
    with pytest.raises(RuntimeError, match='in flight'):
        ss[0].submit_host(*args[0], n_streams=n_streams)
    ss[0].wait_host()
    with pytest.raises(AssertionError):
        ss[0].wait_host()
    ss[1].submit_host(*args[1], n_streams=n_streams)
    for s in ss:
        s.close()          # a context destroyed with a step in flight drains it first


def test_membership_needs_no_scratch_and_is_repeatable():
    """The membership call owns no global scratch: the same call twice into a poisoned output
    buffer gives the same exact rows (dense SUN-RGBD-like scene, 512 boxes = 16-word rows)."""
    f = synth.make_frame(3, 2, N=4000)
    ref = om.points_in_boxes_all_np(f['points'], f['boxes'], 8)
    P, B = torch.from_numpy(f['points']).cuda()[None], torch.from_numpy(f['boxes']).cuda()[None]
    L = G._lib.load()
    out = torch.empty((1, 4000, G.row_words(512)), dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        out.fill_(-1)
        G._lib.check(L.gga_points_in_boxes_bits(P.data_ptr(), 4, B.data_ptr(), out.data_ptr(), 1, 4000, 512, st))
        assert np.array_equal(G.unpack_bits(out, 512)[0].cpu().numpy(), ref)


def test_loss_scratch_contract():
    """loss_sum needs caller-owned scratch: missing / too small / misaligned scratch is rejected."""
    L = G._lib.load()
    n = 64
    p = torch.rand((n, 4), device='cuda')
    p[:, 2:] += p[:, :2]
    t = p.clone()
    ls = torch.zeros((1,), device='cuda')
    sc = torch.zeros((int(L.gga_loss_scratch_bytes()),), dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    args = (p.data_ptr(), t.data_ptr(), None, 1, None, n, G._lib.LOSS_GIOU, 1e-6, 1.0, None, ls.data_ptr(), None, None)
    assert L.gga_box2d_loss(*args, None, 0, st) == -1
    assert L.gga_box2d_loss(*args, sc.data_ptr(), 128, st) == -1
    assert L.gga_box2d_loss(*args, sc.data_ptr() + 4, sc.numel() - 4, st) == -1
    for _ in range(3):   # the kernel leaves the scratch zeroed: reusable without re-initialisation
        assert L.gga_box2d_loss(*args, sc.data_ptr(), sc.numel(), st) == 0
        torch.cuda.synchronize()
        assert abs(float(ls)) < 1e-6
    assert int(sc[:4].view(torch.int32)) == 0


def test_concurrent_steps_do_not_share_state():
    """Two steps replayed from parallel graph branches and from two host threads at once give
    the loss sums of the sequential runs (each GeometryStep owns its reduction scratch)."""
    import threading
    F, N, M = 2, 3000, 300
    steps, sets, want = [], [], []
    for k in range(4):
        bt = synth.make_batch(2, 50 + 3 * k, F, N=N, M=M)
        t = {n_: torch.from_numpy(np.ascontiguousarray(v)).cuda() for n_, v in bt.items()}
        s = GeometryStep(F, N, M, 'cuda', kind='giou', mode='lidar_direct')
        s.run(t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(F * M))
        torch.cuda.synchronize()
        want.append((float(s.loss_sum), s.bits.clone(), s.grad_boxes.clone()))
        steps.append(s)
        sets.append(t)
    streams = [torch.cuda.Stream() for _ in steps]

    def worker(i):
        with torch.cuda.stream(streams[i]):
            for _ in range(20):
                t = sets[i]
                steps[i].run(t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(F * M))
    th = [threading.Thread(target=worker, args=(i,)) for i in range(len(steps))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    torch.cuda.synchronize()
    for s, (ls, bits, gb) in zip(steps, want):
        assert float(s.loss_sum) == ls and torch.equal(s.bits, bits) and torch.equal(s.grad_boxes, gb)
    # the same four steps as parallel branches of one CUDA graph, replayed several times
    g = torch.cuda.CUDAGraph()
    for s in steps:
        s.loss_sum.zero_()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        for i, s in enumerate(steps):
            streams[i].wait_event(fork)
            with torch.cuda.stream(streams[i]):
                t = sets[i]
                for _ in range(3):
                    s.run(t['points'], t['boxes'], t['lidar2img'], t['target'], t['weight'], float(F * M))
            cur.wait_stream(streams[i])
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    for s, (ls, bits, gb) in zip(steps, want):
        assert float(s.loss_sum) == ls and torch.equal(s.bits, bits) and torch.equal(s.grad_boxes, gb)


@pytest.mark.parametrize('T,N,B', [(1, 777, 1), (33, 1000, 2), (1500, 3000, 1), (5000, 1200, 1), (16, 300, 200)])
def test_shapes_beyond_the_training_one(T, N, B):
    """One box, rows wider than 32 words (T > 1024: fewer points per batch), more frames than SMs."""
    rng = np.random.default_rng(T * 7 + N)
    boxes = np.stack([synth.make_boxes(rng, T, 'sunrgbd' if T > 1000 else 'kitti') for _ in range(B)])
    pts = np.stack([synth.make_points(rng, N, boxes[b], 'sunrgbd' if T > 1000 else 'kitti') for b in range(B)])
    ref = np.stack([om.points_in_boxes_all_np(pts[b], boxes[b], 8) for b in range(B)])
    P, Bx = torch.from_numpy(pts).cuda(), torch.from_numpy(boxes).cuda()
    assert np.array_equal(G.unpack_bits(G.points_in_boxes_bits(P, Bx), T).cpu().numpy(), ref)
    assert np.array_equal(G.points_in_boxes_all(P[..., :3], Bx).cpu().numpy(), ref)
    part = G.points_in_boxes_part(P[..., :3], Bx).cpu().numpy()
    assert np.array_equal(part, np.where(ref.any(2), ref.argmax(2), -1))
