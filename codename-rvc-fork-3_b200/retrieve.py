"""Drop-in for `Pipeline._retrieve_speaker_embeddings` (rvc/infer/pipeline.py:497-507).

    def _retrieve_speaker_embeddings(self, feats, index, big_npy, index_rate): ...

Same signature, same return value (a (1, n, d) float32 tensor on feats' device), but the whole body
-- D2H copy, faiss search, numpy weighting / gather / sum, H2D copy, blend -- is one call into
librvcr that never leaves the GPU.  `big_npy` is accepted and ignored: the gather reads the index's
own list-ordered storage.  The reference returns a NEW tensor and keeps `feats0 = feats.clone()`
taken before the call (pipeline.py:455), so blending in place on `feats` is safe; pass
`inplace=False` to keep the input untouched.
"""
from __future__ import annotations

from .faiss_compat import IndexIVFFlat


def retrieve_speaker_embeddings(feats, index: IndexIVFFlat, big_npy=None, index_rate: float = 0.75,
                                inplace: bool = False):
    """feats: (1, n, d) float32 CUDA tensor (HuBERT / ContentVec last_hidden_state).  Like the reference function it
    returns a NEW tensor and leaves `feats` alone; inplace=True blends into `feats` itself (what `patch_pipeline` installs:
    inside `voice_conversion` the un-blended features were cloned before the call, pipeline.py:455) and raises if that is
    impossible (a non-contiguous view) instead of silently working on a copy."""
    if not isinstance(index, IndexIVFFlat):
        raise TypeError("index must come from codename-rvc-fork-3_b200.read_index / index_factory")
    if inplace:
        if not feats.is_contiguous():
            raise ValueError("inplace=True needs a contiguous tensor")
        f = feats
    else:
        f = feats.clone(memory_format=__import__("torch").contiguous_format)
    index.retrieve_blend_(f, index_rate)
    return f


def convert_features(feats, index, index_rate: float, pitchf=None, protect: float = 0.5):
    """Everything between the feature extractor and `net_g.infer` in `Pipeline.voice_conversion`
    (pipeline.py:455-481) as ONE library call when an index is loaded:

        feats0 = feats.clone()                       # :455   (not needed: feats is read only here)
        feats = self._retrieve_speaker_embeddings()  # :456-461
        feats = F.interpolate(..., scale_factor=2)   # :463-465
        feats0 = F.interpolate(...); pitchff = ...; feats = feats * pitchff + feats0 * (1 - pitchff)   # :468-481

    `pitchf` is the (1, p_len) F0 track cut to p_len (pipeline.py:466,474), or None without pitch guidance.  The protect
    mix is applied only when protect < 0.5 (pipeline.py:475) and then needs p_len == 2n exactly, like the reference's
    broadcast `feats * pitchff.unsqueeze(-1)` does; with protect >= 0.5 the track is not looked at (any length).  With
    `index` None / falsy or index_rate <= 0 (pipeline.py:553) only the post step runs.  Returns a new (1, 2n, d) tensor."""
    from .faiss_compat import upsample_protect
    if pitchf is not None and protect >= 0.5:
        pitchf = None                      # pipeline.py:475: no mix, the F0 track plays no role in this step
    if index and index_rate > 0:
        if not isinstance(index, IndexIVFFlat):
            raise TypeError("index must come from codename-rvc-fork-3_b200.read_index / index_factory")
        return index.retrieve_blend_upsample(feats.contiguous(), index_rate, pitchf=pitchf, protect=protect)
    f = feats.contiguous()
    return upsample_protect(f, f if pitchf is not None else None, pitchf, protect)


def _method(self, feats, index, big_npy, index_rate):
    # inside Pipeline.voice_conversion `feats0 = feats.clone()` was taken before this call (pipeline.py:455) and `feats` is
    # rebound to the return value, so blending in place is safe there and saves a (1, n, d) copy
    return retrieve_speaker_embeddings(feats, index, big_npy, index_rate, inplace=bool(getattr(feats, "is_contiguous", lambda: False)()))


def patch_pipeline(pipeline_cls_or_module, faiss_module_name: str = "faiss", cache: bool = True,
                   lazy_big_npy: bool = True) -> None:
    """Install the fused path on the reference without editing it:

        import rvc.infer.pipeline as P
        patch_pipeline(P)            # P.faiss -> this package, P.Pipeline._retrieve_speaker_embeddings -> fused

    After this `faiss.read_index` / `index.reconstruct_n` at pipeline.py:555-556 build a device-resident
    index (big_npy becomes a lazy placeholder that is never materialised) and pipeline.py:459 calls the
    fused kernel path.  `cache=True` keeps the index resident across `pipeline()` calls -- the reference
    re-reads the file for every converted file / chunk (pipeline.py:553-561, infer.py:291-311); see
    faiss_compat.IndexCache."""
    import sys
    from . import faiss_compat
    pkg = sys.modules[__package__]
    target = pipeline_cls_or_module
    if hasattr(target, "Pipeline"):
        setattr(target, faiss_module_name, pkg)
        target = target.Pipeline
    target._retrieve_speaker_embeddings = _method
    if cache:
        faiss_compat.enable_index_cache(True)
    if lazy_big_npy:
        IndexIVFFlat.lazy_reconstruct = True
