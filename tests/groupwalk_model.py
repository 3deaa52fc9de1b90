"""CPU model (float64, pure Python/numpy) of the group walk's opening criterion -- test
infrastructure, not product code.

It builds the reference-shaped octree (root cube and child assignment of
/root/reference/gravhopper/_jbgrav.c:387-462,:764-772: one particle per leaf, strict `>` octant
choice, child centre = centre +- size/4), takes the targets in depth-first leaf order (= the Morton
order of csrc/tree.cu), and for every 32 consecutive targets walks the tree ONCE with the
conservative form of the reference's test (_jbgrav.c:502) that walk_group_kernel uses: the targets
are cut into two bounding boxes at the largest gap between consecutive targets, and a cell is
accepted only if  size^2 / theta^2 < min over both boxes of dist^2(box, cell centre).  Every target
then sums the monopoles of the group's list.  Used to show on the CPU that the criterion is a
refinement of the reference's (list >= accepted set, error <= the reference tree's).
"""
import sys

import numpy as np


class _Node(object):
    __slots__ = ("c", "size", "child", "p", "count", "m", "mx")

    def __init__(self, c, size):
        self.c = c
        self.size = size
        self.child = [None] * 8
        self.p = -1
        self.count = 0
        self.m = 0.0
        self.mx = np.zeros(3)


def build(x, m, eps):
    mn, mx = x.min(0), x.max(0)
    box = (mx[0] - mn[0]) + eps                      # _jbgrav.c:764-769 (padding quirk included)
    for k in (1, 2):
        if (mx[k] - mn[k]) > box:
            box = (mx[k] - mn[k]) + eps
    root = _Node(0.5 * (mn + mx), box)

    def octant(p, c):                                # _jbgrav.c:441-462
        return int(p[0] > c[0]) | (int(p[1] > c[1]) << 1) | (int(p[2] > c[2]) << 2)

    def sub(n, b):
        if n.child[b] is None:
            off = np.array([0.25 if (b >> k) & 1 else -0.25 for k in range(3)]) * n.size
            n.child[b] = _Node(n.c + off, 0.5 * n.size)
        return n.child[b]

    for i in range(len(m)):                          # _jbgrav.c:387-437
        n = root
        while True:
            if n.count == 0:
                n.p, n.count, n.m, n.mx = i, 1, m[i], m[i] * x[i]
                break
            if n.count == 1:
                old, n.p = n.p, -1
                ch = sub(n, octant(x[old], n.c))
                ch.p, ch.count, ch.m, ch.mx = old, 1, m[old], m[old] * x[old]
            n.count += 1
            n.m += m[i]
            n.mx = n.mx + m[i] * x[i]
            n = sub(n, octant(x[i], n.c))
    return root


def leaf_order(root):
    out = []

    def rec(n):
        if n.count == 1:
            out.append(n.p)
            return
        for ch in n.child:
            if ch is not None:
                rec(ch)
    rec(root)
    return out


def _box(pts):
    lo, hi = pts.min(0), pts.max(0)
    return 0.5 * (lo + hi), 0.5 * (hi - lo)


def _dist2(box, c):
    d = np.maximum(np.abs(c - box[0]) - box[1], 0.0)
    return float(d @ d)


def group_walk(x, m, eps, theta, group=32):
    """Accelerations on all particles and the summed list length (entries x targets)."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    root = build(x, m, eps)
    order = leaf_order(root)
    acc = np.zeros_like(x)
    nlist = 0
    it2 = np.inf if theta == 0 else 1.0 / theta ** 2
    for g0 in range(0, len(order), group):
        idx = order[g0:g0 + group]
        pts = x[idx]
        if len(idx) > 1:
            cut = int(np.argmax(((pts[1:] - pts[:-1]) ** 2).sum(1)))
            A, B = _box(pts[:cut + 1]), _box(pts[cut + 1:])
        else:
            A = B = _box(pts)
        com, mass = [], []

        def rec(n):
            if n.count == 1:
                com.append(x[n.p]); mass.append(m[n.p])
                return
            if n.size * n.size * it2 < min(_dist2(A, n.c), _dist2(B, n.c)):
                com.append(n.mx / n.m); mass.append(n.m)
                return
            for ch in n.child:
                if ch is not None:
                    rec(ch)
        rec(root)
        com, mass = np.array(com), np.array(mass)
        nlist += len(mass) * len(idx)
        for i, t in zip(idx, pts):
            d = com - t
            s = (d * d).sum(1) + eps * eps
            w = np.where(s > 0, mass / np.where(s > 0, s, 1.0) ** 1.5, 0.0)
            acc[i] = (w[:, None] * d).sum(0)
    return acc, nlist
