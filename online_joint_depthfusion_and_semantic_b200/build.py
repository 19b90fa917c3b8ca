"""In-tree build of csrc/libojdf.so (hand-written sm_100a kernels + the C ABI of include/ojdf.h).

    python -m online_joint_depthfusion_and_semantic_b200.build

nvcc cross-compiles without a GPU.  -fmad=false is part of the numerics contract: the
reference rounds every product and sum separately, so the compiler must
never contract a*b+c on its own; the kernels spell out the FMAs they do want.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(CSRC, 'libojdf.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-fmad=false', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'ojdf.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get('NVCC', 'nvcc')
    extra = os.environ.get('OJDF_EXTRA_NVCC_FLAGS', '').split()      # e.g. -DOJDF_SS_PROFILE for the role profile of conv_ss
    cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', SO] + sources()
    subprocess.check_call(cmd)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
