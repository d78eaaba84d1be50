#!/usr/bin/env python
"""Stand-in inputs for the reference's regression goldens.

tests/golden/out-seq1.cfrk and out-seq2.cfrk are verbatim copies of the reference's
test/out-seq1.cfrk and test/out-seq2.cfrk (k=2, nt=12, chunkSize=8192, reference
test/test.sh:13-19).  Their inputs, sample/seq1.fasta and sample/seq2.fasta, are missing from the
reference checkout (.MISSING_LARGE_BLOBS), so this script INVERTS the goldens into FASTA files
that reproduce them:

  at k=2 a row is the edge multiset of a multigraph on {A,C,G,T}; walking the rows from the last
  to the first, subtract from bin 15 the spill owed by the next read, split the remaining edges
  into the minimum number p of trails (Hierholzer with p-1 artificial edges per component) and
  join the trails with 'N' (2 invalid windows per joint -> a spill of 2(p-1) into the row above).

The stand-ins are our construction, not the authors' data: they pin the output format, the bin
order, the spill rule and k=2 counting -- not the parser's behaviour on the original files.
Deterministic; `python make_standins.py OUTDIR` writes seq1.standin.fasta / seq2.standin.fasta.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BASES = "ACGT"


def read_rows(path):
    rows = []
    with open(path, "rb") as f:
        for line in f.read().split(b"\n"):
            toks = line.split()
            rows.append([int(t.split(b":")[1]) for t in toks])
    return rows


def trails_of(counts):
    """counts[4a+b] copies of edge a->b  ->  list of trails (strings) covering every edge once."""
    adj = [[counts[4 * a + b] for b in range(4)] for a in range(4)]
    out_d = [sum(adj[a]) for a in range(4)]
    in_d = [sum(adj[a][b] for a in range(4)) for b in range(4)]
    # weakly connected components over nodes that have edges
    comp = list(range(4))

    def find(x):
        while comp[x] != x:
            x = comp[x]
        return x
    for a in range(4):
        for b in range(4):
            if adj[a][b]:
                comp[find(a)] = find(b)
    trails = []
    for root in sorted({find(v) for v in range(4) if out_d[v] + in_d[v] > 0}):
        nodes = [v for v in range(4) if find(v) == root and out_d[v] + in_d[v] > 0]
        surplus_out = []  # nodes that must start a trail
        surplus_in = []   # nodes that must end a trail
        for v in nodes:
            d = out_d[v] - in_d[v]
            surplus_out += [v] * max(0, d)
            surplus_in += [v] * max(0, -d)
        art = {}  # artificial edges (u, v) -> multiplicity
        if surplus_out:
            start = surplus_out[0]
            # connect every other (end -> start) pair, leaving one start and one end open
            for u, v in zip(surplus_in[1:], surplus_out[1:]):
                art[(u, v)] = art.get((u, v), 0) + 1
        else:
            start = nodes[0]
        # Hierholzer on real + artificial edges, deterministic order (real edges first)
        rem = [[adj[a][b] for b in range(4)] for a in range(4)]
        arem = dict(art)
        stack, path = [(start, False)], []
        while stack:
            v, _ = stack[-1]
            nxt = None
            for b in range(4):
                if rem[v][b]:
                    rem[v][b] -= 1
                    nxt = (b, False)
                    break
            if nxt is None:
                for (u, w), m in sorted(arem.items()):
                    if u == v and m:
                        arem[(u, w)] -= 1
                        nxt = (w, True)
                        break
            if nxt is None:
                path.append(stack.pop())
            else:
                stack.append(nxt)
        path.reverse()  # [(node, arrived_by_artificial_edge)]
        cur = BASES[path[0][0]]
        for node, artificial in path[1:]:
            if artificial:
                trails.append(cur)
                cur = BASES[node]
            else:
                cur += BASES[node]
        trails.append(cur)
    return trails


def invert(rows):
    reads = [None] * len(rows)
    owed = 0  # spill of read i+1 into row i
    for i in range(len(rows) - 1, -1, -1):
        own = list(rows[i])
        own[15] -= owed
        if own[15] < 0:
            raise ValueError(f"row {i}: bin 15 smaller than the spill owed by the next read")
        tr = trails_of(own)
        if not tr:
            reads[i] = "A"   # no windows at all: a read of length 1
            owed = 0
        else:
            reads[i] = "N".join(tr)
            owed = 2 * (len(tr) - 1)
    return reads, owed


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name in ("seq1", "seq2"):
        rows = read_rows(os.path.join(HERE, f"out-{name}.cfrk"))
        reads, lost = invert(rows)
        path = os.path.join(outdir, f"{name}.standin.fasta")
        with open(path, "w") as f:
            for i, r in enumerate(reads):
                f.write(f">{name}_standin_{i}\n{r}\n")
        print(f"{path}: {len(reads)} reads, {sum('N' in r for r in reads)} with N, "
              f"spill of read 0 (lost, as in the reference) = {lost}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_standins"))
