"""Parity helpers shared by tests/, smoke() and bench.py's checker leg.

Bars (north_star / SURVEY.md 8c): the interaction set is BIT-EXACT as a sorted set of keys
(type, id_i, id_j, sub_i, sub_j, swap, ghost); floating-point state is compared at `rtol` relative to the field's
own scale (max |x| over the field), i.e. |a-b| <= rtol * max(1e-300, max|b|), AND per element:
|a_i - b_i| <= rtol * |b_i| + ELEM_FLOOR * rtol * max|b| — the floor (1e-2 of the field-scaled bar, i.e. 1e-12 of the field scale
at rtol = 1e-10) only absorbs the rounding of a component that is the small residual of large cancelling contributions (a net
acceleration of 0.35 left over from contact forces of a few thousand moves by 1e-10 when the summation order changes — measured
between the oracle's serial and OpenMP orders, tests/test_oracle_sim.py); a small-magnitude component that is wrong by more than
that fails even though the field-scaled check alone would let it pass.  `abs_floor` (default 0) is for quantities that are
DIFFERENCES of much larger operands (an overlap dn = |r_i - r_j| - R of 1e-7 between coordinates of order 1 cannot be known better
than an ulp of the coordinates): the caller passes a few ulps of the operand scale.  `elem_floor` replaces ELEM_FLOOR, and
`elementwise=False` drops the per-element bar, where the two sides legitimately sum in different orders over many steps (N bricks
against one domain, 60 steps: measured up to 5e-11 of the field scale on components a tenth of it — tests/mgpu_worker.py)."""
from __future__ import annotations

import numpy as np

KEY_FIELDS = ("type", "id_i", "id_j", "sub_i", "sub_j", "swap", "ghost")


def sort_interactions(rec: np.ndarray) -> np.ndarray:
    order = np.lexsort(tuple(rec[f] for f in reversed(KEY_FIELDS)))
    return rec[order]


def keys_of(rec: np.ndarray) -> np.ndarray:
    return np.stack([rec[f].astype(np.uint64) for f in KEY_FIELDS], axis=1)


ELEM_FLOOR = 1e-2


def field_close(a, b, rtol, what="", elementwise=True, abs_floor=0.0, elem_floor=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return 0.0
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert np.array_equal(fa, fb), f"{what}: non-finite mismatch at {int((fa != fb).sum())} elements"
    a, b = np.where(fb, a, 0.0), np.where(fb, b, 0.0)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    diff = np.abs(a - b)
    err = float(np.max(diff)) / scale
    assert err <= rtol, f"{what}: max scaled error {err:.3e} > {rtol:.1e} (scale {scale:.3e})"
    if elementwise and rtol > 0:
        ef = ELEM_FLOOR if elem_floor is None else elem_floor
        bar = rtol * np.abs(b) + max(ef * rtol * scale, abs_floor)
        bad = diff > bar
        if bad.any():
            i = int(np.argmax(diff / bar))
            raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} elements outside |a-b| <= {rtol:.0e}*|b| + {ef * rtol:.0e}*scale; "
                                 f"worst flat index {i}: {a.flat[i]!r} vs {b.flat[i]!r} (scale {scale:.3e})")
    return err


def compare_interaction_sets(a: np.ndarray, b: np.ndarray, payload_rtol: float | None = 1e-10, elem_floor=None):
    """a, b: arrays of capi.INTERACTION_DTYPE. Keys must match exactly; payloads within payload_rtol."""
    assert len(a) == len(b), f"interaction count differs: {len(a)} vs {len(b)}"
    sa, sb = sort_interactions(a), sort_interactions(b)
    ka, kb = keys_of(sa), keys_of(sb)
    if not np.array_equal(ka, kb):
        bad = np.nonzero((ka != kb).any(axis=1))[0]
        raise AssertionError(f"interaction sets differ at {len(bad)} sorted positions; first: {ka[bad[0]]} vs {kb[bad[0]]}")
    if payload_rtol is not None and len(a):
        field_close(sa["friction"], sb["friction"], payload_rtol, "friction history", elem_floor=elem_floor)
        field_close(sa["moment"], sb["moment"], payload_rtol, "moment history", elem_floor=elem_floor)
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


def canonical_interactions(rec: np.ndarray, sphere: bool) -> np.ndarray:
    """Decomposition-independent view of an interaction set.

    With `symetric: true` spheres, a pair that straddles a sub-domain or periodic boundary is enumerated from BOTH sides
    (nbh_sphere.cu:245-249), each side holding a mirrored history record (friction, moment negated).  Which pairs straddle
    a boundary depends on the decomposition, so sets from different decompositions are compared after (i) orienting every
    sphere pair as id_i < id_j (negating the payload when flipped), (ii) clearing the ghost flag, (iii) dropping duplicates.
    Polyhedron records are owned by the lower-id particle only, so (ii) alone makes them comparable."""
    r = rec.copy()
    if sphere:
        flip = (r["type"] == 0) & (r["id_i"] > r["id_j"])
        a, b = r["id_i"][flip].copy(), r["id_j"][flip].copy()
        r["id_i"][flip], r["id_j"][flip] = b, a
        r["friction"][flip] *= -1.0
        r["moment"][flip] *= -1.0
    r["ghost"] = 0
    r["cell_i"] = r["cell_j"] = 0
    r["p_i"] = r["p_j"] = 0
    r = sort_interactions(r)
    k = keys_of(r)
    keep = np.ones(len(r), dtype=bool)
    if len(r) > 1:
        keep[1:] = (k[1:] != k[:-1]).any(axis=1)
    return r[keep]
