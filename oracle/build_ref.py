"""oracle/build_ref.py -- TEST INFRASTRUCTURE, build container only.

Compiles the reference's two native (Cython) modules
    /root/reference/occuspytial/distributions.pyx   (precision_mvnorm, ensure_sums_to_zero)
    /root/reference/occuspytial/data.pyx            (Data)
from the sources WHERE THEY LIE into ``oracle/_ref/occuspytial/`` (git-ignored).
Nothing is copied into the repository.  The checked-in ``.c`` files of the
reference are Cython-0.29 output and are not used; the ``.pyx`` are re-cythonized
with the Cython installed here.  Link flags follow the reference's build.py:14-19
(``npyrandom`` static library).

The pure-Python part of the reference is imported straight from
``/root/reference`` by ``oracle/ref_loader.py``.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('OCCU_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '_ref')


def build():
    src_pkg = os.path.join(REF, 'occuspytial')
    if not os.path.isdir(src_pkg):
        raise SystemExit(f'{src_pkg} not found: the reference is only available in the build container')
    from Cython.Compiler.Main import compile as cy_compile, CompilationOptions

    gen = os.path.join(OUT, 'gen')
    pkg = os.path.join(OUT, 'occuspytial')
    os.makedirs(gen, exist_ok=True)
    os.makedirs(pkg, exist_ok=True)
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    inc = [np.get_include(), sysconfig.get_paths()['include']]
    npylib = os.path.join(np.get_include(), '..', '..', 'random', 'lib')
    for name, libs in (('distributions', ['-L' + npylib, '-lnpyrandom', '-lm']), ('data', [])):
        c_file = os.path.join(gen, name + '.c')
        opts = CompilationOptions(output_file=c_file, language_level=3)
        res = cy_compile(os.path.join(src_pkg, name + '.pyx'), opts,
                         full_module_name='occuspytial.' + name)
        if res.num_errors:
            raise SystemExit(f'cython failed on {name}.pyx')
        so = os.path.join(pkg, name + ext)
        cmd = (['gcc', '-O2', '-fPIC', '-shared', '-w', '-DNPY_NO_DEPRECATED_API=0']
               + ['-I' + i for i in inc] + [c_file, '-o', so] + libs)
        subprocess.check_call(cmd)
        print('built', so)
    shutil.rmtree(gen, ignore_errors=True)


if __name__ == '__main__':
    build()
    sys.exit(0)
