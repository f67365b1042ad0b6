"""The ctypes mirrors of the C-ABI argument structs (gga_b200/_lib.py) have the layout a C
compiler gives the declarations of include/gga_b200.h: size and the offset of every field,
checked by compiling a small C program with the system gcc (the header is plain C)."""
import ctypes
import os
import re
import subprocess

import pytest

from gga_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'gga_b200.h')


def c_fields(struct_name):
    src = open(HEADER).read()
    body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (struct_name, struct_name), src, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        first, *rest = decl.split(',')
        names.append(re.findall(r'(\w+)\s*$', first)[0])
        names += [r.strip().lstrip('*').strip() for r in rest]
    return names


@pytest.mark.parametrize('cname,mirror', [('gga_box_loss_args', _lib.BoxLossArgs), ('gga_target_args', _lib.TargetArgs)])
def test_struct_layout_matches_the_header(tmp_path, cname, mirror):
    names = c_fields(cname)
    assert names == [f[0] for f in mirror._fields_], 'field order / names differ from the header'
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {',
            f'  printf("%zu\\n", sizeof({cname}));']
    prog += [f'  printf("%zu\\n", offsetof({cname}, {n}));' for n in names]
    prog += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(prog))
    exe = tmp_path / 'layout'
    subprocess.run(['/usr/bin/gcc', '-std=c11', '-Wall', '-Werror', '-o', str(exe), str(src)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(mirror)
    for n, off in zip(names, vals[1:]):
        assert getattr(mirror, n).offset == off, n
