"""Exhaustive interleaving check of the peer-memory halo protocol (g4c_halo_put, include/g4c.h) on a small model.

Each rank runs, per exchange k = 1, 2, ...:   PUT(k)   store its rows into every neighbour's mailbox half (k-1) & 1
                                               PUB(k)   publish k to every neighbour's flag slot (after the stores are fenced)
                                               WAIT(k)  blocked until every neighbour's flag in its own slots is >= k
                                               COPY(k)  copy its own mailbox half (k-1) & 1 behind its own rows
                                               USE(k)   the kernels of the block read those rows (any time before PUT(k+1))
Ranks advance independently; the search visits EVERY interleaving of these atomic steps.  Invariant: whatever a rank copies
out (and later uses) in exchange k is what its neighbours wrote in exchange k.  The same search shows why the first version of
the kernel — neighbours storing straight into the ghost rows, no mailbox halves — was wrong: a fast neighbour's PUT(k+1) can
land between a rank's WAIT(k) and its USE(k)."""
import itertools

import pytest

STEPS = ("PUT", "PUB", "WAIT", "COPY", "USE")


def explore(n_ranks, n_exchanges, halves):
    """Breadth-first search over all interleavings.  Returns the first violation found (a string) or None.
    halves = 2: the shipped protocol (double-buffered mailbox); halves = 0: stores go straight to the ghost rows."""
    nbrs = [[q for q in (r - 1, r + 1) if 0 <= q < n_ranks] for r in range(n_ranks)]
    # state: per rank (k, step index); flags[r][q] = last exchange q published to r; mail[r][h][q] = exchange whose rows of q sit
    # in half h of r's mailbox (halves = 0: one slot, the ghost rows themselves); ghost[r][q] = exchange of q's rows r computes with
    start = (tuple((1, 0) for _ in range(n_ranks)),
             tuple(tuple(0 for _ in range(n_ranks)) for _ in range(n_ranks)),
             tuple(tuple(tuple(0 for _ in range(n_ranks)) for _ in range(max(halves, 1))) for _ in range(n_ranks)),
             tuple(tuple(0 for _ in range(n_ranks)) for _ in range(n_ranks)))
    seen, frontier = {start}, [start]
    while frontier:
        nxt = []
        for pos, flags, mail, ghost in frontier:
            for r in range(n_ranks):
                k, si = pos[r]
                if k > n_exchanges:
                    continue
                step = STEPS[si]
                flags2, mail2, ghost2 = flags, mail, ghost
                h = (k - 1) % halves if halves else 0
                if step == "PUT":
                    m = [list(map(list, x)) for x in mail]
                    g = [list(x) for x in ghost]
                    for q in nbrs[r]:
                        if halves:
                            m[q][h][r] = k
                        else:
                            g[q][r] = k                        # straight into q's ghost rows
                    mail2 = tuple(tuple(map(tuple, x)) for x in m)
                    ghost2 = tuple(map(tuple, g))
                elif step == "PUB":
                    f = [list(x) for x in flags]
                    for q in nbrs[r]:
                        f[q][r] = k
                    flags2 = tuple(map(tuple, f))
                elif step == "WAIT":
                    if any(flags[r][q] < k for q in nbrs[r]):
                        continue                               # blocked
                elif step == "COPY":
                    if halves:
                        g = [list(x) for x in ghost]
                        for q in nbrs[r]:
                            if mail[r][h][q] != k:
                                return f"rank {r} copies exchange {mail[r][h][q]} of rank {q} out of its mailbox in exchange {k}"
                            g[r][q] = k
                        ghost2 = tuple(map(tuple, g))
                elif step == "USE":
                    for q in nbrs[r]:
                        if ghost[r][q] != k:
                            return f"rank {r} computes exchange {k} with rows of exchange {ghost[r][q]} from rank {q}"
                npos = list(pos)
                npos[r] = (k, si + 1) if si + 1 < len(STEPS) else (k + 1, 0)
                state = (tuple(npos), flags2, mail2, ghost2)
                if state not in seen:
                    seen.add(state)
                    nxt.append(state)
        frontier = nxt
    return None


@pytest.mark.parametrize("n_ranks,n_exchanges", [(2, 5), (3, 4), (4, 3)])
def test_double_buffered_mailbox_protocol_holds_under_every_interleaving(n_ranks, n_exchanges):
    assert explore(n_ranks, n_exchanges, halves=2) is None


def test_single_mailbox_would_not_be_enough():
    """One mailbox half: a neighbour that has passed WAIT(k) may PUT(k+1) before this rank has copied exchange k out."""
    assert explore(2, 3, halves=1) is not None


def test_direct_stores_into_ghost_rows_race():
    """The first version of the kernel (profiles/r2q -> DESIGN.md 7): 7e-3 error in one of six hardware cases."""
    msg = explore(2, 3, halves=0)
    assert msg is not None and "computes exchange" in msg
