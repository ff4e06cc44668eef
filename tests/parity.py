"""Parity helpers shared by tests/, smoke() and bench.py's checker leg.

Bars (north_star / SURVEY.md 8c): the interaction set is BIT-EXACT as a sorted set of keys
(type, id_i, id_j, sub_i, sub_j, swap, ghost); floating-point state is compared at `rtol` relative to the field's
own scale (max |x| over the field), i.e. |a-b| <= rtol * max(1e-300, max|b|)."""
from __future__ import annotations

import numpy as np

KEY_FIELDS = ("type", "id_i", "id_j", "sub_i", "sub_j", "swap", "ghost")


def sort_interactions(rec: np.ndarray) -> np.ndarray:
    order = np.lexsort(tuple(rec[f] for f in reversed(KEY_FIELDS)))
    return rec[order]


def keys_of(rec: np.ndarray) -> np.ndarray:
    return np.stack([rec[f].astype(np.uint64) for f in KEY_FIELDS], axis=1)


def field_close(a, b, rtol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return 0.0
    assert np.isfinite(a).all() == np.isfinite(b).all(), f"{what}: non-finite mismatch"
    scale = max(float(np.max(np.abs(b))), 1e-300)
    err = float(np.max(np.abs(a - b))) / scale
    assert err <= rtol, f"{what}: max scaled error {err:.3e} > {rtol:.1e} (scale {scale:.3e})"
    return err


def compare_interaction_sets(a: np.ndarray, b: np.ndarray, payload_rtol: float | None = 1e-10):
    """a, b: arrays of capi.INTERACTION_DTYPE. Keys must match exactly; payloads within payload_rtol."""
    assert len(a) == len(b), f"interaction count differs: {len(a)} vs {len(b)}"
    sa, sb = sort_interactions(a), sort_interactions(b)
    ka, kb = keys_of(sa), keys_of(sb)
    if not np.array_equal(ka, kb):
        bad = np.nonzero((ka != kb).any(axis=1))[0]
        raise AssertionError(f"interaction sets differ at {len(bad)} sorted positions; first: {ka[bad[0]]} vs {kb[bad[0]]}")
    if payload_rtol is not None and len(a):
        field_close(sa["friction"], sb["friction"], payload_rtol, "friction history")
        field_close(sa["moment"], sb["moment"], payload_rtol, "moment history")
    return len(a)


def compare_particles(pa, pb, rtol=1e-10, fields=("r", "v", "vrot", "q", "f", "mom", "arot")):
    assert pa.n == pb.n
    assert np.array_equal(pa.id, pb.id), "download order differs"
    errs = {}
    for f in fields:
        errs[f] = field_close(pa.vec(f), pb.vec(f), rtol, f)
    return errs


def bit_identical(pa, pb, fields=("r", "v", "vrot", "q", "f", "mom", "arot")) -> dict:
    return {f: bool(np.array_equal(pa.vec(f), pb.vec(f))) for f in fields}
