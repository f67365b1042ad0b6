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


def test_a_plain_c_program_can_load_and_call_the_library(tmp_path):
    """dlopen from C (no Python, no torch in the process): the entry points that need no device."""
    lib = _lib.library_path() if hasattr(_lib, 'library_path') else os.path.join(ROOT, 'gga_b200', '_C', 'libgga_b200.so')
    if not os.path.isfile(lib):
        from gga_b200 import build
        build.build(verbose=False)
    prog = r'''
#include <dlfcn.h>
#include <stdio.h>
#include "%s"
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "%%s\n", dlerror()); return 2; }
  int (*version)(void) = (int (*)(void))dlsym(h, "gga_version");
  int (*row_words)(int) = (int (*)(int))dlsym(h, "gga_pib_row_words");
  size_t (*ws_bytes)(void) = (size_t (*)(void))dlsym(h, "gga_loss_scratch_bytes");
  int (*bits)(const float*, int, const float*, uint32_t*, int, int, int, void*) =
      (int (*)(const float*, int, const float*, uint32_t*, int, int, int, void*))dlsym(h, "gga_points_in_boxes_bits");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "gga_last_error");
  if (!version || !row_words || !ws_bytes || !bits || !last_error) return 3;
  int rc = bits(NULL, 2, NULL, NULL, 1, 8, 8, NULL);   /* pts_stride < 3: rejected before any CUDA call */
  printf("%%d %%d %%d %%zu %%d %%s\n", version(), row_words(256), row_words(1024), ws_bytes(), rc, last_error());
  return 0;
}
''' % HEADER
    src = tmp_path / 'use.c'
    src.write_text(prog)
    exe = tmp_path / 'use'
    subprocess.run(['/usr/bin/gcc', '-std=c11', '-Wall', '-o', str(exe), str(src), '-ldl'], check=True)
    out = subprocess.run([str(exe), lib], check=True, capture_output=True, text=True).stdout.split(None, 5)
    assert int(out[0]) >= 100 and int(out[1]) == 8 and int(out[2]) == 32 and int(out[3]) > 0
    assert int(out[4]) == -1 and 'pts_stride' in out[5]
