"""
ORACLE — TEST INFRASTRUCTURE ONLY.

Compiles the C restatements under oracle/ into oracle/_build/liboracle.so with gcc.
`-ffp-contract=off` keeps  a + b*c  as two rounded operations, as CPython evaluates it.
The reference itself is pure Python (no compilable sources), so there is no oracle/_ref.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'liboracle.so')
SOURCES = ['occgrid_oracle.c', 'merge_oracle.c']


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs)):
        return LIB
    cmd = ['gcc', '-O2', '-std=c11', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared',
           '-Wall', '-Wextra', '-o', LIB] + srcs + ['-lm']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
