"""The comparator networks of the sparse kernels (cfrk_b200/csrc/sparse.cu sort_lane / merge_lane), read from the
source and checked with the 0-1 principle: the lane networks must sort every 0-1 input, the mergers every input that
can reach them -- the rotations of ascending sequences (tools/host/lane_merge_search.py derives that set from the
cross-lane stages of half_sort)."""
import itertools
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "cfrk_b200", "csrc", "sparse.cu")).read()


def networks(fn):
    """{E: [(i, j), ...]} from the body of `fn` (if constexpr (E == 9) {...} else {...})"""
    body = SRC[SRC.index(f"void {fn}(KeyT (&k)[E])"):]
    body = body[:body.index("#undef CFRK_CE")]
    first, second = body.split("} else {")
    ce = lambda t: [(int(a), int(b)) for a, b in re.findall(r"CFRK_CE\((\d+), (\d+)\)", t.split("#define")[-1].split("\n", 1)[1])]
    return {9: ce(first), 11: ce(second)}


def run(net, p):
    p = list(p)
    for i, j in net:
        if p[i] > p[j]:
            p[i], p[j] = p[j], p[i]
    return p


@pytest.mark.parametrize("E", [9, 11])
def test_lane_network_sorts_everything(E):
    net = networks("sort_lane")[E]
    assert len(net) == {9: 25, 11: 35}[E]
    assert all(0 <= i < j < E for i, j in net)
    for p in itertools.product((0, 1), repeat=E):
        assert run(net, p) == sorted(p)


@pytest.mark.parametrize("E", [9, 11])
def test_lane_merger_sorts_every_bitonic_sequence(E):
    net = networks("merge_lane")[E]
    assert len(net) == {9: 18, 11: 25}[E]
    assert all(0 <= i < j < E for i, j in net)
    for ones in range(E + 1):
        base = [0] * (E - ones) + [1] * ones
        for r in range(E):
            p = base[r:] + base[:r]
            assert run(net, p) == sorted(p), p
    # and not everything: it is a merger, cheaper than a sorting network
    assert any(run(net, p) != sorted(p) for p in itertools.product((0, 1), repeat=E))


def test_reachable_patterns_are_the_rotations():
    """the cross-lane stages of half_sort (mirrored compare + lane strides), simulated on every pair of sorted 0-1 runs
    of the 2-lane and 4-lane levels (E = 9): a lane ends up with a rotation of an ascending sequence, never anything else"""
    E = 9
    rot = {tuple(([0] * (E - o) + [1] * o)[r:] + ([0] * (E - o) + [1] * o)[:r]) for o in range(E + 1) for r in range(E)}
    for m in (2, 4):
        half = (m // 2) * E
        for za in range(half + 1):
            for zb in range(half + 1):
                key = [[0] * E for _ in range(m)]
                for idx in range(half):
                    key[idx // E][idx % E] = 0 if idx < za else 1
                    key[m // 2 + idx // E][idx % E] = 0 if idx < zb else 1
                new = [row[:] for row in key]
                for hl in range(m):
                    upper = (hl & (m >> 1)) != 0
                    for e in range(E):
                        a, b = key[hl][e], key[hl ^ (m - 1)][E - 1 - e]
                        new[hl][e] = max(a, b) if upper else min(a, b)
                key, st = new, m >> 2
                while st >= 1:
                    new = [row[:] for row in key]
                    for hl in range(m):
                        upper = (hl & st) != 0
                        for e in range(E):
                            a, b = key[hl][e], key[hl ^ st][e]
                            new[hl][e] = max(a, b) if upper else min(a, b)
                    key, st = new, st >> 1
                assert all(tuple(r) in rot for r in key)
                flat = [x for r in key for x in sorted(r)]
                assert flat == sorted(flat)
