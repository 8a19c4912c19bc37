"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (kimkyeonghun/MSA) for validating the oracle
restatement and for generating golden vectors.  Never imported by the product (msa_b200/).

The reference is located through $MSA_REF (default /root/reference).  It is driven exactly as
SURVEY.md Appendix A documents, with two harness-side patches and no edits to the reference:
  * transformers-5.x shim: MMBertModel.__init__ calls the legacy ``init_weights()`` without
    ``post_init()`` (MMBertForPretraining.py:22); route the first call through ``post_init``.
  * hidden != 1024: ``MMBertEmbedding.TEXTDIM`` is read as a module global (MMBertEmbedding.py:48-52)
    and ``CPC(x_size=1024)`` is hard-coded (MMBertForPretraining.py:327-344).
"""
import os
import sys

REF_DIR = os.environ.get("MSA_REF", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "MMBertForPretraining.py"))


_mods = None


def load_modules():
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference not found under {REF_DIR}")
    from transformers import PreTrainedModel
    if not getattr(PreTrainedModel, "_mmb_shimmed", False):
        _orig = PreTrainedModel.init_weights

        def _iw(self):
            if "all_tied_weights_keys" not in self.__dict__:
                return self.post_init()
            return _orig(self)

        PreTrainedModel.init_weights = _iw
        PreTrainedModel._mmb_shimmed = True
    # The reference's modules are top-level names (config, MMBertEmbedding, MMBertForPretraining) and
    # transformers resolves ``sys.modules[cls.__module__]``, so they must stay registered under those names.
    # The product never imports these top-level names (its own root-level shims re-export msa_b200.api and
    # are only exercised in a subprocess by the tests), so there is no clash inside one process.
    for k in ("config", "MMBertEmbedding", "MMBertForPretraining"):
        mod = sys.modules.get(k)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF_DIR):
            del sys.modules[k]
    sys.path.insert(0, REF_DIR)
    try:
        import MMBertForPretraining as M
        import MMBertEmbedding as E
        ref_mods = {k: sys.modules[k] for k in ("config", "MMBertEmbedding", "MMBertForPretraining")}
    finally:
        sys.path.remove(REF_DIR)
    _mods = (M, E, ref_mods)
    return _mods


def build_model(cfg, dataset="mosi", eager=True):
    """Constructs the reference MMBertForPretraining for a transformers BertConfig."""
    M, E, _ = load_modules()
    if eager:
        cfg._attn_implementation = "eager"
    E.TEXTDIM = cfg.hidden_size
    model = M.MMBertForPretraining(cfg)
    model.bert.set_joint_embeddings(dataset)
    if cfg.hidden_size != 1024:
        for k in ("cpc_zt", "cpc_zv", "cpc_za"):
            setattr(model, k, E.CPC(x_size=cfg.hidden_size, y_size=cfg.hidden_size))
    return model
