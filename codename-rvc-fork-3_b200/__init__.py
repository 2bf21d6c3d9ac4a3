"""codename-rvc-fork-3_b200 -- B200-native RVC feature retrieval (faiss IVF-Flat search k=8 ->
gather -> 1/score^2 weighting -> index_rate blend) behind the reference's own Python call sites.

The directory name is not a Python identifier; import it with
    importlib.import_module("codename-rvc-fork-3_b200")      (repo root on sys.path)
or through the `rvcr` alias module at the repo root.

Host side only: every number is computed by librvcr.so (hand-written sm_100a CUDA, include/rvcr.h).
"""
from ._ffi import RvcrError, device_count, lib  # noqa: F401
from .faiss_compat import (METRIC_L2, IndexCache, IndexIVFFlat, LazyBigNpy, MiniBatchKMeans, enable_index_cache,  # noqa: F401
                           extract_index_ivf, final_proj, index_cache, index_factory, probe_index, read_index,
                           shard_range, upsample_protect, write_index)
from . import sharded, synth  # noqa: F401
from .retrieve import convert_features, patch_pipeline, retrieve_speaker_embeddings  # noqa: F401

__all__ = ["IndexIVFFlat", "read_index", "write_index", "index_factory", "extract_index_ivf", "METRIC_L2",
           "retrieve_speaker_embeddings", "patch_pipeline", "enable_index_cache", "index_cache", "IndexCache",
           "LazyBigNpy", "upsample_protect", "final_proj", "MiniBatchKMeans", "probe_index", "shard_range", "convert_features", "RvcrError", "device_count", "lib"]
