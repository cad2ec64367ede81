"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the snapshot).

    python tools/install_reference.py [--force]

pip builds from a copy under /tmp (/root/reference is read-only and setuptools writes build/ + egg-info into the source
tree); --no-deps because the wheelhouse has no torch wheel to "resolve" (torch is already in the image).  The reference's
scripts/ and YAML configs are not package data, so they are copied next to the package (baseline/_ref/scripts,
baseline/_ref/fact_clip/configs/*.yaml): tests/test_scripts_unchanged.py runs them as they are.
Nothing under baseline/_ref is tracked by git; nothing in the product imports it.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF, DST = '/root/reference', os.path.join(ROOT, 'baseline', '_ref')


def install(force=False):
    if os.path.isdir(os.path.join(DST, 'fact_clip')) and os.path.isdir(os.path.join(DST, 'scripts')) and not force:
        return DST
    if not os.path.isdir(REF):
        return None
    tmp = '/tmp/_factclip_ref_src'
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp)
    shutil.rmtree(DST, ignore_errors=True)
    subprocess.check_call([sys.executable, '-m', 'pip', 'install', '-q', '--no-index', '--no-build-isolation', '--no-deps',
                           '--find-links', '/opt/wheelhouse', '--target', DST, tmp])
    shutil.copytree(os.path.join(REF, 'scripts'), os.path.join(DST, 'scripts'))
    for f in os.listdir(os.path.join(REF, 'fact_clip', 'configs')):
        if f.endswith(('.yaml', '.yml')):
            shutil.copy(os.path.join(REF, 'fact_clip', 'configs', f), os.path.join(DST, 'fact_clip', 'configs', f))
    shutil.rmtree(tmp, ignore_errors=True)
    return DST


if __name__ == '__main__':
    print(install(force='--force' in sys.argv))
