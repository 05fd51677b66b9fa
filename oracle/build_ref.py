"""Recipe for ``oracle/_ref``: the reference's own implementation of the hot path, made able to travel.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/origin_oracle.py``): nothing under ``origin_b200/``
imports this.

The reference (``musevlt/origin``) is pure Python: there is nothing to compile.  Its hot path lives
in ONE file, ``muse_origin/lib_origin.py``, which only needs numpy / scipy / joblib once the imports
of astropy, mpdaf, photutils and matplotlib are stubbed (``oracle/ref_loader.py``); the step layer that
calls it is ``muse_origin/steps.py`` (its ``Step`` / ``DataObj`` machinery runs with the same stubs).  ``/root/reference``
exists in the build container only, so ``__graft_entry__.build()`` calls :func:`build_ref` there: the
file is copied UNMODIFIED into the git-ignored ``oracle/_ref/muse_origin/`` (never committed — no
reference source enters the history), next to a ``PROVENANCE`` note with its sha256.  ``oracle/_ref``
is not gpurun-ignored, so the copy travels to the GPU box like the built ``libogn.so`` and
``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own functions there
(``kind: "reference"``).  When neither ``oracle/_ref`` nor ``/root/reference`` is present the
bench falls back to the oracle port and says ``kind: "port"``.

    python -m oracle.build_ref
"""

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
REFERENCE_ROOT = os.environ.get('ORIGIN_REFERENCE_ROOT', '/root/reference')
FILES = ['muse_origin/lib_origin.py',      # the hot path: CPU baseline of bench.py, oracle of steps 04 / 08
         'muse_origin/steps.py']           # the step layer: the drop-in tests run the reference's own Step classes


def build_ref(force=False):
    """Copy the reference's hot-path module(s) into ``oracle/_ref``.  Returns the directory, or None when
    the reference tree is not mounted (GPU box: the prebuilt copy is used as is)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return REF_DIR if os.path.isdir(REF_DIR) else None
    lines = []
    for rel in FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
        with open(dst, 'rb') as f:
            lines.append('%s  %s  (unmodified copy of %s)' % (hashlib.sha256(f.read()).hexdigest(), rel, src))
    with open(os.path.join(REF_DIR, 'PROVENANCE'), 'w') as f:
        f.write('Build artefact of oracle/build_ref.py; git-ignored; never edit.\n' + '\n'.join(lines) + '\n')
    return REF_DIR


if __name__ == '__main__':
    print(build_ref(force=True))
