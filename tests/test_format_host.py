"""Host-side logic of gga_b200.kitti_format without a GPU: the KITTI text writer and the annos
rewrite, against the reference's own outputs (tests/golden/ref_format.npz).  The GPU matcher is
replaced by a numpy restatement of image_box_overlap + argmax (eval.py:85-114,
utils_pseudo_labels_gga.py:60) for this test only."""
import copy
import os

import numpy as np
import torch

from gga_b200 import kitti_format as KF
from gga_b200 import synth
from oracle.gen_golden import ANNO_KEYS, FORMAT_COUNTS

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_format.npz'))


def golden_annos():
    return [{k: GOLD[f'f{f}_{k}'] for k in ANNO_KEYS} for f in range(len(FORMAT_COUNTS))]


def test_kitti_lines_reproduce_the_reference_text():
    for f, a in enumerate(golden_annos()):
        assert KF.kitti_lines(a) == str(GOLD[f'f{f}_txt'])


def numpy_matcher(dt, do, gt, go, return_overlaps=False):
    dt, gt, m = dt.cpu().numpy(), gt.cpu().numpy(), []
    for f in range(len(do) - 1):
        g = gt[int(go[f]):int(go[f + 1])]
        for d in dt[int(do[f]):int(do[f + 1])]:
            iw = np.minimum(d[2], g[:, 2]) - np.maximum(d[0], g[:, 0])
            ih = np.minimum(d[3], g[:, 3]) - np.maximum(d[1], g[:, 1])
            inter = np.where((iw > 0) & (ih > 0), iw * ih, 0.0)
            ua = (d[2] - d[0]) * (d[3] - d[1]) + (g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1]) - inter
            m.append(int(np.argmax(inter / ua)))
    return torch.tensor(m, dtype=torch.int32), None


def test_annos_rewrite_equals_reference(monkeypatch):
    monkeypatch.setattr(KF, 'match_dt_to_gt', numpy_matcher)
    infos, _ = synth.make_detection_frames(2024, FORMAT_COUNTS)
    cleaned, new_infos = KF.pseudo_label_matching_kitti(copy.deepcopy(infos), golden_annos(), 0, 200, out_path=None, device='cpu',
                                                        return_infos=True)
    for f in range(len(infos)):
        ref_keys = [k[len(f'f{f}_new_'):] for k in GOLD.files if k.startswith(f'f{f}_new_')]
        assert list(new_infos[f]['annos'].keys()) == ref_keys
        for k in ref_keys:
            ref, got = GOLD[f'f{f}_new_{k}'], np.asarray(new_infos[f]['annos'][k])
            assert got.shape == ref.shape
            assert list(got) == list(ref) if got.dtype.kind in 'US' else np.array_equal(got, ref), (f, k)
        for k in [k[len(f'f{f}_clean_'):] for k in GOLD.files if k.startswith(f'f{f}_clean_')]:
            ref, got = GOLD[f'f{f}_clean_{k}'], np.asarray(cleaned[f][k])
            assert list(got) == list(ref) if got.dtype.kind in 'US' else np.array_equal(got, ref), (f, k)


def test_bbox2result_kitti_needs_the_gpu():
    infos, dets = synth.make_detection_frames(1, (3,))
    if not torch.cuda.is_available():
        try:
            KF.bbox2result_kitti(dets, ['Pedestrian', 'Cyclist', 'Car'], data_infos=infos,
                                 pcd_limit_range=list(synth.KITTI_MATCH_RANGE), device='cpu')
        except AssertionError as e:
            assert 'no CPU path' in str(e)
        else:
            raise AssertionError('bbox2result_kitti must not run without a CUDA device')
