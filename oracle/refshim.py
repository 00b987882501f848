"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference modules.

Only usable in the build container where ``/root/reference`` is mounted (the GPU
box does not have it).  Used by ``oracle/make_golden.py`` to pin the oracle
restatement against the reference's own ``JETSGenerator`` and to generate the
committed fixtures under ``tests/golden/``.  Nothing in the product path
(``emotivoice_b200``) may import this module.

The reference builds its config with ``yacs`` (absent in this image,
inference_am_vocoder_joint.py:53-58); the attr-dict below stands in for it.
"""
import os
import sys

REF_ROOT = os.environ.get("EMOTIVOICE_REFERENCE", "/root/reference")


class AttrDict(dict):
    """yaml -> attribute access.  __getattr__ must raise AttributeError (not
    KeyError) or copy/pickle of ``Generator.h`` breaks."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(o):
    if isinstance(o, dict):
        return AttrDict({k: _wrap(v) for k, v in o.items()})
    return o


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "prompt_tts_modified"))


def load_reference_config(n_vocab=502, n_speaker=2014):
    import yaml
    with open(os.path.join(REF_ROOT, "config", "joint", "config.yaml")) as f:
        conf = _wrap(yaml.safe_load(f))
    conf.n_vocab = n_vocab          # inference_am_vocoder_joint.py:57
    conf.n_speaker = n_speaker      # inference_am_vocoder_joint.py:58
    return conf


def import_reference_jets():
    """Returns the reference's JETSGenerator class (jets.py:26)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from models.prompt_tts_modified.jets import JETSGenerator  # noqa
    return JETSGenerator
