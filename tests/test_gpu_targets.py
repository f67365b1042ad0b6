"""CUDA target packing (gga_pack_targets through gga_b200.targets) vs the reference's own
get_targets_single outputs (tests/golden/ref_targets.npz) and vs the oracle on larger batches.
Bit-exact: indices, masks, heatmaps (fp32 bit patterns), annotations."""
import os

import numpy as np
import pytest
import torch

import gga_b200 as G
from gga_b200 import synth
from gga_b200 import targets as T
from oracle import targets as ot
from oracle.gen_golden import TARGET_CASES

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_targets.npz'))
FRAME_KEYS = ('labels', 'boxes_img', 'lidar2img', 'pseudo', 'bdry', 'base_lidar2img')


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize('case', TARGET_CASES, ids=[c[0] for c in TARGET_CASES])
def test_get_targets_equals_reference_outputs(case):
    name, class_names, over, counts, dt = case
    cfg = dict(synth.KITTI_TRAIN_CFG, **over)
    frames = [{k: GOLD[f'{name}_f{f}_{k}'] for k in FRAME_KEYS} for f in range(len(counts))]
    srl = torch.from_numpy(np.stack([GOLD[f'{name}_f{f}_srl'] for f in range(len(counts))]))
    ibp = [[torch.full((1 + i % 3, 4), float(i), dtype=torch.float64) for i in range(n)] for n in counts]
    out = T.get_targets([torch.from_numpy(fr['labels']).cuda() for fr in frames],
                        [torch.from_numpy(fr['boxes_img']) for fr in frames],
                        [torch.from_numpy(fr['lidar2img']) for fr in frames],
                        [torch.from_numpy(fr['pseudo']) for fr in frames],
                        [torch.from_numpy(fr['bdry']) for fr in frames], ibp,
                        [dict(lidar2img=fr['base_lidar2img']) for fr in frames], class_names, cfg, srl=srl)
    heat, anno, ind, mask, l2i, tibp, bm = out
    K = cfg['max_objs'] * cfg['dense_reg']
    for t in range(len(class_names)):
        assert ind[t].dtype == torch.int64 and mask[t].dtype == torch.uint8 and bm[t].dtype == torch.uint8
        for f in range(len(counts)):
            g = lambda k: GOLD[f'{name}_f{f}_t{t}_{k}']  # noqa: E731
            assert np.array_equal(ind[t][f].cpu().numpy(), g('ind'))
            assert np.array_equal(mask[t][f].cpu().numpy(), g('mask'))
            assert np.array_equal(bits(heat[t][f].cpu().numpy()), bits(g('heatmap')))
            assert np.array_equal(bits(anno[t][f].cpu().numpy()), bits(g('anno_box')))
            assert np.array_equal(bits(l2i[t][f].cpu().numpy()), bits(g('lidar2img')))
            assert np.array_equal(bm[t][f].cpu().numpy(), g('bmask'))
            assert [int(p[0, 0]) for p in tibp[t][f]] == list(g('ibp_order')[:K])
            assert all(p.is_cuda for p in tibp[t][f])


def test_seeded_semantic_ratio_samples_reproduce_the_reference_draws():
    name = TARGET_CASES[0][0]
    torch.manual_seed(1000)
    s = T.semantic_ratio_samples(1, 3)[0].numpy()
    assert np.array_equal(s, GOLD[f'{name}_f0_srl'])
    anno = GOLD[f'{name}_f0_t2_anno_box']
    m = GOLD[f'{name}_f0_t2_mask'].astype(bool)
    assert m.any() and np.all(anno[m, 4] == s[2])


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_training_batch_equals_oracle(dt):
    """8 frames x up to 500 objects (the reference's max_objs), KITTI tasks."""
    rng = np.random.default_rng(77)
    counts = [500, 333, 0, 1, 480, 256, 77, 499]
    frames = [synth.make_target_frame(rng, n, 3, dt) for n in counts]
    srl = rng.uniform(0.5, 4, (len(counts), 3)).astype(np.float32)
    fo = np.concatenate([[0], np.cumsum(counts)])
    cat = lambda k: np.concatenate([fr[k] for fr in frames], 0)  # noqa: E731
    p = T.pack_targets(torch.from_numpy(cat('labels')), fo, torch.from_numpy(cat('boxes_img')),
                       torch.from_numpy(cat('lidar2img')), torch.from_numpy(cat('pseudo')),
                       torch.from_numpy(cat('bdry')), np.stack([fr['base_lidar2img'] for fr in frames]), srl,
                       synth.KITTI_TASKS, synth.KITTI_TRAIN_CFG, device='cuda')
    n_valid = 0
    for f, fr in enumerate(frames):
        heat, anno, ind, mask, l2i, src, bm = ot.get_targets_single(
            fr['labels'], fr['boxes_img'], fr['lidar2img'], fr['pseudo'], fr['bdry'], fr['base_lidar2img'], srl[f],
            synth.KITTI_TASKS, synth.KITTI_TRAIN_CFG)
        for t in range(3):
            assert np.array_equal(p.ind[t, f].cpu().numpy(), ind[t])
            assert np.array_equal(p.mask[t, f].cpu().numpy(), mask[t])
            assert np.array_equal(bits(p.heatmap[f, t].cpu().numpy()), bits(heat[t][0]))
            assert np.array_equal(bits(p.anno_box[t, f].cpu().numpy()), bits(anno[t]))
            assert np.array_equal(bits(p.anno_lidar2img[t, f].cpu().numpy()), bits(l2i[t]))
            assert np.array_equal(p.boundary_mask[t, f].cpu().numpy(), bm[t])
            s = p.src_index[t, f].cpu().numpy()
            assert np.array_equal(np.where(s >= 0, s - fo[f], -1), src[t])
            n_valid += int(mask[t].sum())
    assert n_valid > 1500


def test_empty_batch_and_bad_arguments():
    cfg = synth.KITTI_TRAIN_CFG
    e = torch.zeros
    p = T.pack_targets(e((0,), dtype=torch.int64), [0, 0], e((0, 4)), e((0, 4, 4)), e((0, 7), dtype=torch.float64),
                       e((0, 4), dtype=torch.bool), np.eye(4, dtype=np.float32)[None], np.ones((1, 3), np.float32),
                       synth.KITTI_TASKS, cfg, device='cuda')
    assert p.heatmap.abs().sum().item() == 0 and p.mask.sum().item() == 0 and (p.src_index == -1).all()
    assert torch.equal(p.anno_lidar2img[1, 0, 7].cpu(), torch.eye(4))
    rng = np.random.default_rng(0)
    fr = synth.make_target_frame(rng, 3000, 3, np.float32, adversarial=False)
    with pytest.raises(RuntimeError, match='objects per frame'):
        T.pack_targets(torch.from_numpy(fr['labels']), [0, 3000], torch.from_numpy(fr['boxes_img']),
                       torch.from_numpy(fr['lidar2img']), torch.from_numpy(fr['pseudo']), torch.from_numpy(fr['bdry']),
                       fr['base_lidar2img'][None], np.ones((1, 3), np.float32), synth.KITTI_TASKS, cfg, device='cuda')
