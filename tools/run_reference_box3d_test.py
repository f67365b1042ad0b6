"""Drop-in acceptance: runs the REFERENCE'S OWN test `tests/test_utils/test_box3d.py::test_points_in_boxes`
(CUDA-gated, test_box3d.py:1683-1797) through the reference's own box classes (base_box3d.py:510-580,
cam_box3d.py:303-354) with this library's CUDA ops injected as `mmcv.ops.points_in_boxes_all/_part`.

Needs the reference tree (GGA_REFERENCE_ROOT, default /root/reference) and a GPU.  Test
infrastructure: the reference files are loaded read-only by oracle/ref_loader.py; nothing here is on
the product path.  Usage:  python tools/run_reference_box3d_test.py [extra pytest args]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import pytest
    import torch
    import gga_b200 as G
    from oracle import ref_loader
    calls = {'all': 0, 'part': 0}

    def pib_all(points, boxes):
        calls['all'] += 1
        return G.points_in_boxes_all(points, boxes)

    def pib_part(points, boxes):
        calls['part'] += 1
        return G.points_in_boxes_part(points, boxes)

    ref_loader.load_reference(pib_all, pib_part)
    ref_loader.expose_for_reference_tests()
    test = os.path.join(ref_loader.REF_ROOT, 'tests/test_utils/test_box3d.py')
    print('CUDA available:', torch.cuda.is_available(), '| library:', G._lib.lib_path())
    rc = pytest.main(['-q', '-p', 'no:cacheprovider', '--rootdir', os.path.dirname(test), f'{test}::test_points_in_boxes',
                      '-rs'] + sys.argv[1:])
    print(f'injected op calls: points_in_boxes_all x{calls["all"]}, points_in_boxes_part x{calls["part"]}')
    if torch.cuda.is_available():
        assert calls['all'] > 0 and calls['part'] > 0, 'the injected CUDA ops were never called'
    return int(rc)


if __name__ == '__main__':
    sys.exit(main())
