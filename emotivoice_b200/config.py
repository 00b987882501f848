"""Model hyper-parameters of the hot path.

Mirrors the keys the reference reads from ``config/joint/config.yaml:36-94``
(``model.*``) and the attributes its callers patch onto the yacs node
(``n_vocab`` / ``n_speaker``: inference_am_vocoder_joint.py:57-58;
``n_mels`` / ``segment_size``: config.yaml).  Any attr-object exposing the same
names (a yacs ``CfgNode``, the reference's own config) is accepted by the
modules in :mod:`emotivoice_b200.modules`; this file only provides a default so
the engine is usable without the reference tree.
"""


class AttrDict(dict):
    """dict with attribute access.  ``__getattr__`` raises AttributeError (not
    KeyError) so that ``copy.deepcopy`` / pickle work."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        import copy
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def wrap(o):
    if isinstance(o, dict) and not isinstance(o, AttrDict):
        return AttrDict({k: wrap(v) for k, v in o.items()})
    return o


_MODEL_DEFAULT = dict(
    speaker_embed_dim=384, bert_embedding=768,
    encoder_n_layers=4, encoder_n_heads=8, encoder_n_hidden=384,
    encoder_p_dropout=0.2, encoder_kernel_size_conv_mod=3,
    decoder_n_layers=4, decoder_n_heads=8, decoder_n_hidden=384,
    decoder_p_dropout=0.2, decoder_kernel_size_conv_mod=3,
    variance_n_hidden=384, variance_n_layers=3, variance_kernel_size=3,
    variance_p_dropout=0.1, variance_embed_kernel_size=9,
    variance_embde_p_dropout=0.0,
    duration_p_dropout=0.5, duration_n_layers=2, duration_kernel_size=3,
    resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
    initial_channel=80, upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
)


def default_config(n_vocab=502, n_speaker=2014):
    """The released EmotiVoice joint model (config.yaml:36-94; 502 phoneme
    symbols in data/youdao/text/tokenlist, 2014 speakers in .../speaker2)."""
    return AttrDict(
        model=AttrDict({k: (list(v) if isinstance(v, list) else v) for k, v in _MODEL_DEFAULT.items()}),
        n_mels=80, segment_size=32, n_vocab=n_vocab, n_speaker=n_speaker,
        sr=16000, hop_length=256,
    )


def load_yaml_config(path, n_vocab=502, n_speaker=2014):
    import yaml
    with open(path) as f:
        conf = wrap(yaml.safe_load(f))
    conf.n_vocab = n_vocab
    conf.n_speaker = n_speaker
    return conf
