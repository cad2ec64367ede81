"""Build libfactk.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m fact_clip_b200.build [--force]

The shared library is written next to this file so it travels with the repo snapshot to the GPU box.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libfactk.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def build(force=False, verbose=False):
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(HERE, '..', 'include', 'factk.h')]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, 'build', os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) >= _newest([s] + deps[len(srcs):]):
            continue
        procs.append((s, subprocess.Popen([nvcc] + NVCC_FLAGS + ['-c', s, '-o', o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- nvcc {os.path.basename(s)}\n{out}\n')
        if p.returncode != 0:
            failed = True
        else:
            with open(os.path.join(HERE, 'build', os.path.basename(s)[:-3] + '.ptxas.log'), 'w') as f:
                f.write(out)
    if failed:
        raise RuntimeError('nvcc failed')
    subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-lcuda', '-lcudart'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
