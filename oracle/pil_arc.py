"""ORACLE (test infrastructure only): restatement of Pillow's `ImageDraw.arc` rasteriser.

The reference draws its "spaghetti" with PIL (/root/reference/argus/utils.py:252-275: `d.arc((x0, y0, x1, y1), start, end,
fill, width=int(width))`). Pillow (>= 8.0; pinned here against the installed Pillow by tests/test_oracle_augment.py, which
draws thousands of random arcs with the real library and demands IDENTICAL pixels) rasterises an arc as:

  1. an integer Bresenham-style walk over one quadrant of the outer ellipse (semi-axes a = x1 - x0, b = y1 - y0 in
     DOUBLED coordinates: a point (X, Y) of the walk is pixel (x0 + (X + a) / 2, y0 + (Y + b) / 2)) and of the inner
     ellipse (a - 2 (w - 1), b - 2 (w - 1)); every row gets the horizontal segments between the two,
  2. clipped by two half planes: the NORMAL lines of the ellipse at the parametric points of the start / end angle
     (AND of the two when the sweep is below 180 degrees, OR otherwise), with the wide / tall cases made symmetric by a
     transposition.
"""
from __future__ import annotations

import math

import numpy as np


class _Quarter:
    """One quadrant of the ellipse with semi-axes (a, b) in doubled coordinates, walked from (a, b % 2) to (a % 2, b)."""

    def __init__(self, a: int, b: int) -> None:
        self.finished = a < 0 or b < 0
        if self.finished:
            return
        self.a, self.b = a, b
        self.cx, self.cy = a, b % 2
        self.ex, self.ey = a % 2, b
        self.a2, self.b2 = a * a, b * b
        self.a2b2 = self.a2 * self.b2

    def _delta(self, x: int, y: int) -> int:
        return abs(self.a2 * y * y + self.b2 * x * x - self.a2b2)

    def next(self):
        if self.finished:
            return None
        ret = (self.cx, self.cy)
        if self.cx == self.ex and self.cy == self.ey:
            self.finished = True
        else:
            nx, ny = self.cx, self.cy + 2
            nd = self._delta(nx, ny)
            if nx > 1:
                d = self._delta(self.cx - 2, self.cy + 2)
                if nd > d:
                    nx, ny, nd = self.cx - 2, self.cy + 2, d
                d = self._delta(self.cx - 2, self.cy)
                if nd > d:
                    nx, ny = self.cx - 2, self.cy
            self.cx, self.cy = nx, ny
        return ret


def ellipse_rows(a: int, b: int, w: int) -> dict[int, tuple[int, int]]:
    """{Y >= 0: (l, r)} in doubled coordinates: on rows +-Y the ring covers X in [l, r] and [-r, -l]
    (a single segment [-r, r] when l <= 0)."""
    rows: dict[int, tuple[int, int]] = {}
    outer = _Quarter(a, b)
    first = outer.next()
    if w < 1 or first is None:
        return rows
    inner = _Quarter(a - 2 * (w - 1), b - 2 * (w - 1))
    leftmost = a % 2
    pr, py = first
    pl = leftmost
    finished = False
    while not finished:
        y, l, r = py, pl, pr
        while True:
            nxt = outer.next()
            if nxt is None or nxt[1] > y:
                break
        if nxt is None:
            finished = True
        else:
            pr, py = nxt
        while True:
            nxt = inner.next()
            if nxt is None or nxt[1] > y:
                break
            l = nxt[0]
        pl = leftmost if nxt is None else nxt[0]
        rows[y] = (l, r)
    return rows


def _round_up(f: float) -> int:
    """Pillow's ROUND_UP: nearest integer, ties away from zero."""
    return int(math.floor(f + 0.5)) if f >= 0 else -int(math.floor(abs(f) + 0.5))


def _round_down(f: float) -> int:
    """Pillow's ROUND_DOWN: nearest integer, ties towards zero."""
    return int(math.ceil(f - 0.5)) if f >= 0 else -int(math.ceil(abs(f) - 0.5))


_INF = 1 << 30


def _halfplane(node, y: int):
    """Integer interval of the scan coordinate x with A x + B y + C >= 0 (None = empty)."""
    A, B, C = node
    eps = 1e-9
    if A > eps:
        return (_round_up(-(B * y + C) / A), _INF)
    if A < -eps:
        return (-_INF, _round_down(-(B * y + C) / A))
    return (-_INF, _INF) if B * y + C >= -eps else None


def _isect(i, j):
    if i is None or j is None:
        return None
    lo, hi = max(i[0], j[0]), min(i[1], j[1])
    return (lo, hi) if lo <= hi else None


def normalize_angles(al: float, ar: float) -> tuple[float, float]:
    if ar - al >= 360:
        return 0.0, 360.0
    al = math.fmod(al, 360)
    if al < 0:
        al += 360
    d = math.fmod(ar - al, 360)
    if d < 0:
        d += 360
    return al, al + d


def half_rules(al: float, ar: float) -> list[str]:
    """What the arc [al, ar] (degrees, al in [0, 360), ar in (al, al + 360)) leaves of the half ellipse k = 0
    (angles 0..180, Y >= 0 on a y-down screen) and k = 1 (180..360): 'none', 'all', 'nl' (from the start cap on),
    'nr' (up to the end cap), 'and' (between the caps) or 'or' (everything but the gap between the caps)."""
    rules = []
    e = ar if ar < 360 else ar - 360
    for k in range(2):
        q0, q1 = 180.0 * k, 180.0 * k + 180.0
        has_start = q0 <= al < q1
        has_end = q0 < e <= q1
        if has_start and has_end:
            rules.append("and" if (ar - al) < 180 else "or")
        elif has_start:
            rules.append("nl")
        elif has_end:
            rules.append("nr")
        else:
            mid = (q0 + q1) / 2
            rules.append("all" if (al <= mid <= ar) or (al <= mid + 360 <= ar) else "none")
    return rules


def arc_row_intervals(a: int, b: int, w: int, start: float, end: float) -> dict[int, list[tuple[int, int]]]:
    """{pixel row offset (0..b): [(first, last) pixel column offsets (0..a), ...]} of `ImageDraw.arc((x0, y0, x0 + a,
    y0 + b), start, end, width=w)`; offsets are relative to (x0, y0). Intervals may overlap."""
    out: dict[int, list[tuple[int, int]]] = {}
    al, ar = normalize_angles(float(start), float(end))
    if ar == al or a < 0 or b < 0:
        return out
    full = ar == al + 360
    rows = ellipse_rows(a, b, w)
    if not full:
        # the clip tree is built for a WIDE ellipse (a normal line of a wide ellipse leaves it in the other vertical
        # half, so "Y >= 0" / "Y <= 0" nodes are enough to keep the two caps apart); a tall one is handled in the
        # transposed frame and the tree is transposed back
        transposed = a < b
        if transposed:
            A_, B_ = b, a
            al2, ar2 = normalize_angles(90 - ar, 90 - al)
        else:
            A_, B_, al2, ar2 = a, b, al, ar
        lc = (-A_ * math.sin(al2 * math.pi / 180.0), B_ * math.cos(al2 * math.pi / 180.0),
              (A_ * A_ - B_ * B_) * math.sin(al2 * math.pi / 90.0) / 2.0)
        rc = (A_ * math.sin(ar2 * math.pi / 180.0), -B_ * math.cos(ar2 * math.pi / 180.0),
              (B_ * B_ - A_ * A_) * math.sin(ar2 * math.pi / 90.0) / 2.0)
        rules = half_rules(al2, ar2)
        halves = [(0.0, 1.0, 0.0), (0.0, -1.0, 0.0)]
        if transposed:
            lc, rc = (lc[1], lc[0], lc[2]), (rc[1], rc[0], rc[2])
            halves = [(n[1], n[0], n[2]) for n in halves]
    for Y, (l, r) in rows.items():
        for sy in ((Y, -Y) if Y > 0 else (Y,)):
            if not (l > 0 or l < r):
                segs = [(-r, r)]
            else:
                # solid row of an even-width ellipse (l == 0): the centre pixel belongs to the left segment only
                segs = [(-r, -l), (l if l > 0 else 2, r)]
            ivs = []
            for seg in segs:
                if full:
                    ivs.append(seg)
                    continue
                hl, hr = _halfplane(lc, sy), _halfplane(rc, sy)
                for k in range(2):
                    base = _isect(seg, _halfplane(halves[k], sy))
                    rule = rules[k]
                    if rule == "none" or base is None:
                        continue
                    if rule == "all":
                        ivs.append(base)
                    elif rule == "nl":
                        ivs.append(_isect(base, hl))
                    elif rule == "nr":
                        ivs.append(_isect(base, hr))
                    elif rule == "and":
                        ivs.append(_isect(_isect(base, hl), hr))
                    else:
                        ivs.append(_isect(base, hl))
                        ivs.append(_isect(base, hr))
            px = [((o[0] + a) // 2, (o[1] + a) // 2) for o in ivs if o is not None]
            px = [(max(p0, 0), min(p1, a)) for p0, p1 in px if p1 >= 0 and p0 <= a and p0 <= p1]
            if px:
                out.setdefault((sy + b) // 2, []).extend(px)
    return out


def arc_mask(H: int, W: int, bbox: tuple[int, int, int, int], start: float, end: float, width: int) -> np.ndarray:
    """Boolean (H, W) mask of the pixels `ImageDraw.Draw(img).arc(bbox, start, end, fill, width)` paints."""
    x0, y0, x1, y1 = (int(v) for v in bbox)
    m = np.zeros((H, W), dtype=bool)
    for row, ivs in arc_row_intervals(x1 - x0, y1 - y0, int(width), start, end).items():
        y = y0 + row
        if 0 <= y < H:
            for p0, p1 in ivs:
                m[y, max(x0 + p0, 0):min(x0 + p1, W - 1) + 1] = True
    return m
