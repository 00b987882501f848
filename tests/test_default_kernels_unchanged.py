"""The device code of the DEFAULT path must be exactly the code that last passed on the B200.

After round 1's GPU budget ended the kernels were templated on launch-mode flags (EV_PDL) and the library gained opt-in
kernels.  `profiles/r01_sass_default_kernel_hashes.json` holds the per-kernel hashes of the SASS instruction streams of the
flag-off instantiations; they were checked to be byte-identical to the GPU-validated revision (683b582 + the same nvcc), and
this test keeps them that way.  After a deliberate, GPU-validated kernel change: `python tools/sass_hashes.py --write <file>`."""
import json
import os
import shutil
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump (CUDA toolkit)")
def test_default_path_kernels_are_the_gpu_validated_ones():
    from emotivoice_b200 import build
    build.build(verbose=False)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_hashes
    with open(os.path.join(ROOT, "profiles", "r01_sass_default_kernel_hashes.json")) as f:
        want = json.load(f)
    got = sass_hashes.kernel_hashes()
    assert sorted(got) == sorted(want), "kernel set changed: %s" % sorted(set(got) ^ set(want))
    changed = [k for k in want if got[k] != want[k]]
    assert not changed, "device code of default-path kernels changed (re-validate on the GPU, then rewrite the hash file): %s" % changed
