"""The bit arithmetic of octree.cuh's search bookkeeping, restated in Python and checked exhaustively (the CUDA code itself is
exercised by the octree-mode parity tests on the GPU; this pins the FORMULAS a reader finds there).

oct_order_bytes:       per octant one byte = 1 << (its position in the search order), octants 0..3 in .x, 4..7 in .y
oct_present_in_order:  child mask (bit i = octant i has a child) -> the same set as bits in SEARCH order: every mask bit is spread to
                       a byte (nibble * 0x00204081 puts bit k at 8k without carries), selects its octant's byte, bytes OR-ed together
child index:           first child + popcount(mask below bit i)   (children of a parent are neighbours, in octant order)
"""
import itertools
import random

M32 = 0xFFFFFFFF


def order_bytes(order):
    bx = by = 0
    for p in range(8):
        i = (order >> (3 * p)) & 7
        v = ((1 << p) << (8 * (i & 3))) & M32
        if i & 4:
            by |= v
        else:
            bx |= v
    return bx, by


def present_in_order(mask, b):
    mlo = ((((mask & 15) * 0x00204081) & 0x01010101) * 0xFF) & M32
    mhi = (((((mask >> 4) & 15) * 0x00204081) & 0x01010101) * 0xFF) & M32
    x = (b[0] & mlo) | (b[1] & mhi)
    x |= x >> 16
    x |= x >> 8
    return x & 0xFF


def test_nibble_spread_has_no_carries():
    for n in range(16):
        spread = ((n * 0x00204081) & 0x01010101) & M32
        assert spread == sum(((n >> k) & 1) << (8 * k) for k in range(4))


def test_present_children_in_search_order():
    rng = random.Random(1)
    perms = [list(p) for p in itertools.islice(itertools.permutations(range(8)), 0, None, 97)]   # every 97th of the 40 320 orders
    perms += [rng.sample(range(8), 8) for _ in range(200)]
    for perm in perms:
        order = sum(perm[p] << (3 * p) for p in range(8))
        b = order_bytes(order)
        for mask in range(256):
            want = sum(((mask >> perm[p]) & 1) << p for p in range(8))
            assert present_in_order(mask, b) == want


def test_walking_the_bits_visits_children_in_search_order():
    rng = random.Random(2)
    for _ in range(300):
        perm = rng.sample(range(8), 8)
        order = sum(perm[p] << (3 * p) for p in range(8))
        mask = rng.randrange(1, 256)
        rem = present_in_order(mask, order_bytes(order))
        visited = []
        while rem:
            p = (rem & -rem).bit_length() - 1          # __ffs(rem) - 1
            rem &= rem - 1
            i = (order >> (3 * p)) & 7
            visited.append((i, bin(mask & ((1 << i) - 1)).count("1")))   # octant, offset from the first child
        assert [v[0] for v in visited] == [o for o in perm if (mask >> o) & 1]
        assert sorted(v[1] for v in visited) == list(range(bin(mask).count("1")))
