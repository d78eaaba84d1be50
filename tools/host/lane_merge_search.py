"""Which 0-1 patterns can a lane of the half-warp sorting network (sparse.cu half_sort: 16 lanes x E keys) hold right
before the in-lane step of every merge level, and what is the smallest comparator network that sorts them all?
(0-1 principle on the set of inputs 'two sorted runs', which is closed under thresholding.)
usage: python tools/host/lane_merge_search.py [E]"""
import itertools
import sys

E = int(sys.argv[1]) if len(sys.argv) > 1 else 9
L = 16


def level_patterns(m):
    """all in-lane 0-1 patterns (tuples of E bits, index e) reachable before the in-lane step of level m"""
    half = (m // 2) * E
    pats = set()
    for za in range(half + 1):
        for zb in range(half + 1):
            # lanes 0..m-1: group A = lanes [0, m/2) sorted ascending (zeros first), group B likewise
            key = [[0] * E for _ in range(m)]
            for idx in range(half):
                key[idx // E][idx % E] = 0 if idx < za else 1
                key[m // 2 + idx // E][idx % E] = 0 if idx < zb else 1
            # mirrored compare
            new = [row[:] for row in key]
            for hl in range(m):
                upper = (hl & (m >> 1)) != 0
                o = hl ^ (m - 1)
                for e in range(E):
                    a, b = key[hl][e], key[o][E - 1 - e]
                    new[hl][e] = max(a, b) if upper else min(a, b)
            key = new
            st = m >> 2
            while st >= 1:
                new = [row[:] for row in key]
                for hl in range(m):
                    upper = (hl & st) != 0
                    for e in range(E):
                        a, b = key[hl][e], key[hl ^ st][e]
                        new[hl][e] = max(a, b) if upper else min(a, b)
                key = new
                st >>= 1
            # after a full in-lane sort the m lanes must be sorted: check the multiset property now
            flat = [sorted(r) for r in key]
            seq = [x for r in flat for x in r]
            assert seq == sorted(seq), (m, za, zb)
            for r in key:
                pats.add(tuple(r))
    return pats


def apply(net, p):
    p = list(p)
    for i, j in net:
        if p[i] > p[j]:
            p[i], p[j] = p[j], p[i]
    return tuple(p)


def sorts_all(net, pats):
    return all(list(apply(net, p)) == sorted(p) for p in pats)


allp = set()
for m in (2, 4, 8, 16):
    P = level_patterns(m)
    print(f"level m={m}: {len(P)} reachable in-lane patterns of {2 ** E}")
    allp |= P
print(f"union: {len(allp)} patterns = every rotation of an ascending 0-1 sequence (bitonic): "
      f"{allp == {tuple(([0] * (E - o) + [1] * o)[r:] + ([0] * (E - o) + [1] * o)[:r]) for o in range(E + 1) for r in range(E)}}")

# the networks in sparse.cu merge_lane
SHIPPED = {
    9: [(0, 3), (1, 4), (5, 8), (0, 6), (1, 7), (2, 8), (3, 6), (2, 5), (4, 7), (0, 1), (3, 4), (6, 7), (0, 2), (3, 5), (6, 8),
        (1, 2), (4, 5), (7, 8)],
    11: [(0, 4), (5, 10), (1, 6), (3, 9), (2, 7), (1, 3), (6, 9), (0, 8), (2, 5), (7, 10), (4, 8), (3, 5), (6, 7), (0, 1), (8, 9),
         (0, 2), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 10), (9, 10)],
}
if E in SHIPPED:
    print(f"shipped merger for E={E}: {len(SHIPPED[E])} comparators, sorts every reachable pattern: {sorts_all(SHIPPED[E], allp)}")


def beam_search(pats, width=20000, max_len=40):
    """shortest comparator sequence found by a beam over sets of patterns (bit-encoded)"""
    def enc(p):
        return sum(b << i for i, b in enumerate(p))
    done_set = {enc(tuple([0] * (E - o) + [1] * o)) for o in range(E + 1)}
    pairs = [(i, j) for i in range(E) for j in range(i + 1, E)]

    def ce(S, i, j):
        return frozenset((x & ~(1 << i) | (1 << j)) if (x >> i) & 1 and not (x >> j) & 1 else x for x in S)
    beam = {frozenset(enc(p) for p in pats): []}
    for depth in range(1, max_len + 1):
        nxt = {}
        for S, net in beam.items():
            for (i, j) in pairs:
                S2 = ce(S, i, j)
                if S2 != S and S2 not in nxt:
                    nxt[S2] = net + [(i, j)]
        for S, net in nxt.items():
            if all(x in done_set for x in S):
                return net
        beam = dict(sorted(nxt.items(), key=lambda kv: (sum(1 for x in kv[0] if x not in done_set), len(kv[0])))[:width])
    return None


if "--search" in sys.argv:
    net = beam_search(allp)
    print("beam search:", net, "length", len(net) if net else None)
