"""Copy the UNMODIFIED reference (suryanshkumar/online-joint-depthfusion-and-semantic) into baseline/_ref/.

    python baseline/install_ref.py            # from /root/reference (or $OJDF_REFERENCE)

The reference is pure Python with no setup.py, so "installing" it is copying its Python tree (modules/, utils/,
dataset/, the four drivers, configs/, lists/; deps/ -- offline data preparation, 9 MB of Cython/C++ -- is left
out).  baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so it travels to the GPU box, where `bench.py --impl reference` and the drive-through tests run the
reference's own code on the host CPU (baseline/harness.py supplies the stand-ins for the packages the image lacks).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
KEEP = ('modules', 'utils', 'dataset', 'configs', 'lists', 'test_fusion.py', 'train_fusion.py', 'test_segmentation.py',
        'train_segmentation.py', 'LICENSE.md', 'README.md', 'environment.yml')


def install(src=None, quiet=False):
    src = src or os.environ.get('OJDF_REFERENCE', '/root/reference')
    if not os.path.isdir(os.path.join(src, 'modules')):
        if not quiet:
            print('no reference tree at %s: nothing installed' % src)
        return None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for name in KEEP:
        s = os.path.join(src, name)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(DST, name), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        elif os.path.isfile(s):
            shutil.copy2(s, os.path.join(DST, name))
    if not quiet:
        print('reference installed into %s' % DST)
    return DST


if __name__ == '__main__':
    sys.exit(0 if install() else 1)
