"""Host-side pieces of gga_b200.targets that need no GPU."""
import os

import numpy as np
import pytest
import torch

from gga_b200 import synth
from gga_b200 import targets as T
from oracle.gen_golden import TARGET_CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_targets.npz'))


def test_seeded_semantic_ratio_samples_equal_the_reference_draws():
    """torch.manual_seed + the reference's draw order (centerpoint_head_gga.py:515-527) reproduce
    the srl column the reference's get_targets_single wrote into anno_box."""
    name, class_names = TARGET_CASES[0][0], TARGET_CASES[0][1]
    for f in range(3):
        torch.manual_seed(1000 + f)
        s = T.semantic_ratio_samples(1, len(class_names))[0].numpy()
        assert np.array_equal(s, GOLD[f'{name}_f{f}_srl'])
        for t in range(len(class_names)):
            m = GOLD[f'{name}_f{f}_t{t}_mask'].astype(bool)
            assert np.all(GOLD[f'{name}_f{f}_t{t}_anno_box'][m, 4] == s[t])
    assert (T.semantic_ratio_samples(4, 3) >= 1e-3).all()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_pack_targets_has_no_cpu_path():
    fr = synth.make_target_frame(np.random.default_rng(0), 5, 3, np.float32, adversarial=False)
    with pytest.raises(AssertionError, match='no CPU path'):
        T.pack_targets(torch.from_numpy(fr['labels']), [0, 5], torch.from_numpy(fr['boxes_img']),
                       torch.from_numpy(fr['lidar2img']), torch.from_numpy(fr['pseudo']), torch.from_numpy(fr['bdry']),
                       fr['base_lidar2img'][None], np.ones((1, 3), np.float32), synth.KITTI_TASKS, synth.KITTI_TRAIN_CFG,
                       device='cpu')
