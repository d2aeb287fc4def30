"""Results of a bulk (multi-GPU) pre-quantization pass, keyed by initializer name.

The reference's rewriter calls the plugin one node at a time from a single thread
(qrules/_common.py:133), so sharding weights over ranks has to happen *before* the rewrite: a
pre-pass quantizes every target weight (``parallel.shard``), stores the triples here, and the
registered plugins' ``quantize_weights`` return the stored triple instead of recomputing it.
"""
from __future__ import annotations

_store: dict[str, tuple] = {}


def put(name: str, triple: tuple) -> None:
    _store[name] = triple


def lookup(w):
    name = getattr(w, "name", None)
    return _store.get(name) if name is not None else None


def clear() -> None:
    _store.clear()
