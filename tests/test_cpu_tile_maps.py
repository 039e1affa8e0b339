"""Index maps of the shared-memory tile (csrc/tisph_lists.cuh, csrc/tisph_kernels.cuh), restated in Python and
checked without a GPU: the candidate <-> slot permutation is a bijection on every aligned run of 16, its 16
(lane, parity) classes cut a lattice neighbourhood more evenly than the plain map, and the incremental range search of the staging loops
(tile_to_global_fwd) names the same global row as the counting form (tile_to_global) for every candidate of
every thread -- including empty ranges, ragged tiles and the second record of the density staging."""
import random

NB_THREADS = 256


def cand_to_slot(e):            # tisph_lists.cuh: cand_to_slot
    x, y, z = e >> 4, e >> 2, e
    return (e & ~15) | (((x - y) & 3) << 2) | ((x + y + z) & 3)


def slot_to_cand(s):            # tisph_lists.cuh: slot_to_cand
    x, chi, clo = s >> 4, s >> 2, s
    y = (x - chi) & 3
    return (s & ~15) | (y << 2) | ((clo - x - y) & 3)


def tile_to_global(off, gb, e):             # tisph_kernels.cuh: tile_to_global
    k = sum(1 for t in range(1, 9) if e >= off[t])
    return gb[k] + (e - off[k])


def tile_to_global_fwd(off, gb, e, k):      # tisph_kernels.cuh: tile_to_global_fwd (k is a one-element list)
    while k[0] < 8 and e >= off[k[0] + 1]:
        k[0] += 1
    return gb[k[0]] + (e - off[k[0]])


def test_candidate_slot_permutation_is_a_bijection_inside_every_run_of_16():
    for base in range(0, 2048, 16):
        slots = [cand_to_slot(e) for e in range(base, base + 16)]
        assert sorted(slots) == list(range(base, base + 16))
        assert all(slot_to_cand(cand_to_slot(e)) == e for e in range(base, base + 16))


def test_lattice_neighbourhood_is_spread_more_evenly_than_by_the_plain_map():
    """a cell of the reference lattice is 4 x 4 x 4 candidates (x = bits 4-5, y = bits 2-3, z = bits 0-1 of the
    candidate).  The plain map slot = candidate hands class k (slot mod 16) the column (y, z) = (k >> 2, k & 3),
    which a ball of neighbours either contains or misses: lists of 4 and of 0 entries side by side.  The
    permuted classes are diagonals of the sub-lattice: no class is left empty, none takes a whole column."""
    for centre in ((1.5, 1.5, 1.5), (0.5, 1.5, 2.5)):
        perm, plain = [0] * 16, [0] * 16
        for e in range(64):
            x, y, z = (e >> 4) & 3, (e >> 2) & 3, e & 3
            if (x - centre[0]) ** 2 + (y - centre[1]) ** 2 + (z - centre[2]) ** 2 <= 2.8:
                perm[cand_to_slot(e) & 15] += 1
                plain[e & 15] += 1
        assert sum(perm) == sum(plain)
        assert min(plain) == 0 and min(perm) >= 1
        assert max(perm) - min(perm) < max(plain) - min(plain)


def _random_ranges(rng):
    while True:
        lens = [rng.choice([0, 0, 1, 5, 63, 64, 130, 192, 200, 250]) for _ in range(9)]
        if 0 < sum(lens) <= 2032:
            break
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    gb = [rng.randrange(0, 10 ** 6) for _ in range(9)]
    return off, gb


def test_incremental_range_search_matches_the_counting_form_in_both_staging_loops():
    rng = random.Random(7)
    for _ in range(300):
        off, gb = _random_ranges(rng)
        total = off[9]
        # force walk: thread tid stages candidates tid, tid + 256, ...
        want = {cand_to_slot(e): tile_to_global(off, gb, e) for e in range(total)}
        got = {}
        for tid in range(NB_THREADS):
            k = [0]
            for e in range(tid, total, NB_THREADS):
                got[cand_to_slot(e)] = tile_to_global_fwd(off, gb, e, k)
        assert got == want
        # density walk: records p = tid, tid + 256, ... of two slots (rows 2k, 2k + 1) each, two records per
        # iteration; a ghost cell that skips the walk stages nothing (walk_total = 0)
        for walk_total in (total, 0):
            nrow2 = ((walk_total + 15) // 16 + 7) // 8 * 8
            n8 = nrow2 * 8
            want = {}
            for p in range(n8):
                sa = 16 * (p >> 3) + (p & 7)
                for s in (sa, sa + 8):
                    e = slot_to_cand(s)
                    want[s] = tile_to_global(off, gb, e) if e < walk_total else None
            got = {}
            for tid in range(NB_THREADS):
                ka, kb = [0], [0]
                for p in range(tid, n8, 2 * NB_THREADS):
                    two = p + NB_THREADS < n8
                    for u in range(2):
                        pu = p + u * NB_THREADS
                        sa = 16 * (pu >> 3) + (pu & 7)
                        ea, eb = slot_to_cand(sa), slot_to_cand(sa + 8)
                        va = (u == 0 or two) and ea < walk_total
                        vb = (u == 0 or two) and eb < walk_total
                        ga = tile_to_global_fwd(off, gb, ea, ka) if va else None
                        gb_ = tile_to_global_fwd(off, gb, eb, kb) if vb else None
                        if u == 0 or two:
                            assert sa not in got and sa + 8 not in got
                            got[sa], got[sa + 8] = ga, gb_
            assert got == want
