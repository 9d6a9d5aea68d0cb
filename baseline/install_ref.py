#!/usr/bin/env python
"""Stage the UNMODIFIED reference where the bench's reference arm and the tier-A oracle can import it on the GPU box.

The contract's recipe -- ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference`` -- cannot work: Ariostgx/ProSim ships neither setup.py nor pyproject.toml ("Directory is not installable";
outcome recorded in DESIGN.md).  The reference is a plain source tree that is used by putting it on sys.path
(README.md of the reference: ``pip install -e`` is never mentioned; scripts run from the checkout), so "installing" it is
copying its Python package: this script copies ``prosim/**/*.py`` and the yaml configs verbatim into ``baseline/_ref/``
(git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so).  No source file is edited; the two 20 MB
scene-id tables (``*.pkl``, read at import, unused by the rollout) are replaced by empty pickles, the demo assets stay behind.  ``oracle/ref_shim.py`` then imports it exactly as it imports /root/reference.

    python baseline/install_ref.py [/root/reference]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')


def install(src='/root/reference', verbose=True):
    if not os.path.isdir(os.path.join(src, 'prosim')):
        if verbose:
            print(f'install_ref: no reference tree at {src}; keeping whatever is in {DST}')
        return os.path.isdir(os.path.join(DST, 'prosim'))
    pip = subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--find-links',
                          '/opt/wheelhouse', '--target', DST + '_pip', src], capture_output=True, text=True)
    note = (pip.stderr.strip().splitlines() or ['ok'])[-1]
    shutil.rmtree(DST + '_pip', ignore_errors=True)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for top in ('prosim', 'prosim_demo/cfg'):
        for root, _, files in os.walk(os.path.join(src, top)):
            for f in files:
                if f.endswith(('.py', '.yaml', '.yml')):
                    rel = os.path.relpath(os.path.join(root, f), src)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel))
                    n += 1
    # prosim/dataset/data_utils.py:57-61 unpickles two scene-id tables (20 MB, ProSim-Instruct-520k labels) at import time;
    # the rollout path never reads them: empty stand-ins keep the import working without shipping them to every GPU box
    import pickle
    lab = os.path.join(DST, 'prosim', 'dataset', 'prosim_instruct_520k')
    os.makedirs(lab, exist_ok=True)
    for split in ('train', 'val'):
        with open(os.path.join(lab, f'waymo_{split}_IDs.pkl'), 'wb') as fh:
            pickle.dump({}, fh)
    with open(os.path.join(DST, 'INSTALL_NOTE.txt'), 'w') as fh:
        fh.write(f'pip install outcome: {note}\ncopied {n} files verbatim from {src}\n')
    if verbose:
        print(f'install_ref: pip said "{note}"; staged {n} reference files in {DST}')
    return True


if __name__ == '__main__':
    install(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
