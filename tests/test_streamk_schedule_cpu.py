"""CPU model of the stream-K work assignment of the TMA ZGEMM (``seg_count`` / ``get_seg`` and the host rule that sets
``streamk`` / ``sk_tiles`` / the grid in pytdscf_b200/csrc/zgemm_tma.cu), checked for the properties the kernel relies on:
every (tile, k-tile) is computed exactly once; a tile cut over several CTAs is published by CTA c and continued by CTA c + 1
(never a CTA with a lower index waiting for a higher one: in-order CTA dispatch then makes the wait deadlock-free); a CTA
publishes at most once per launch (one partial buffer and one flag per CTA) and does so BEFORE anything it waits for.
The model restates the C++ line by line; the GPU tests (tests/test_gpu_bench_shapes.py) pin the kernel itself."""
import pytest

G_SMS = 148


def host_rule(total: int, KT: int, G: int = G_SMS, force_tma: bool = False):
    """-> (streamk, sk_tiles, grid): zgemm_tma_try, the block that follows ``p.streamk = 0``."""
    grid = min(total, G)
    streamk, sk_tiles = 0, total
    if KT >= 16:
        if total >= G:
            waves = (total + G - 1) // G
            streamk = 1 if ((waves * G - total) * 100 > 3 * waves * G or force_tma) else 0
            if total % G == 0:
                streamk = 0
            sk_tiles = total % G + (G if total >= 2 * G else 0)
            if total < 2 * G:
                sk_tiles = total
        else:
            sk_tiles = total
            if 2 * total >= G and KT >= 32:
                streamk, grid = 1, G
            elif force_tma:
                g = min(total * KT // 4, G)
                if g > total:
                    streamk, grid = 1, g
    return streamk, sk_tiles, grid


def segments(cta: int, grid: int, total: int, KT: int, streamk: int, sk_tiles: int):
    """-> [(tile, kt0, kt1, role)] in execution order: get_seg for si = 0 .. seg_count - 1."""
    if not streamk:
        return [(w, 0, KT, 0) for w in range(cta, total, grid)]
    I = sk_tiles * KT
    a, b = I * cta // grid, I * (cta + 1) // grid
    out = []
    if b > a:
        first, a_off, last = a // KT, a % KT, (b - 1) // KT
        b_off = b - last * KT
        if first == last:
            role = 0 if (a_off == 0 and b_off == KT) else (1 if a_off == 0 else (2 if b_off == KT else 3))
            out.append((first, a_off, b_off, role))
        else:
            n = last - first + 1
            for si in range(n):
                if si == 0:
                    out.append((last, 0, b_off, 0 if b_off == KT else 1))
                elif si == n - 1:
                    out.append((first, a_off, KT, 0 if a_off == 0 else 2))
                else:
                    out.append((first + si, 0, KT, 0))
    rest = total - sk_tiles
    assert rest % grid == 0
    out += [(sk_tiles + w * grid + cta, 0, KT, 0) for w in range(rest // grid)]
    return out


CASES = [(t, kt, f) for f in (False, True) for kt in (16, 17, 32, 64, 96, 512) for t in
         (1, 8, 40, 73, 74, 80, 96, 147, 148, 149, 150, 216, 295, 296, 297, 384, 512, 1024, 4096, 4097)]


@pytest.mark.parametrize("total,KT,force", CASES)
def test_streamk_assignment_properties(total, KT, force):
    streamk, sk_tiles, grid = host_rule(total, KT, force_tma=force)
    assert 1 <= grid <= G_SMS
    covered = {}
    publisher, continuer = {}, {}
    for cta in range(grid):
        segs = segments(cta, grid, total, KT, streamk, sk_tiles)
        publishes = [s for s in segs if s[3] in (1, 3)]
        assert len(publishes) <= 1                                  # one partial buffer / flag per CTA and launch
        waited = False
        for tile, k0, k1, role in segs:
            assert 0 <= tile < total and 0 <= k0 < k1 <= KT
            for k in range(k0, k1):
                assert (tile, k) not in covered, (tile, k)
                covered[(tile, k)] = cta
            if role in (2, 3):
                waited = True
                continuer.setdefault(tile, []).append((cta, k0, k1))
            if role == 1:
                assert not waited                                    # a head is published before this CTA waits for anything
            if role in (1, 3):
                publisher.setdefault(tile, []).append((cta, k0, k1))
            if role == 1:
                assert k0 == 0
            if role == 2:
                assert k1 == KT
    assert len(covered) == total * KT                                # everything exactly once
    for tile, cont in continuer.items():
        for cta, k0, _ in cont:                                      # the CTA that continues a tile reads its predecessor's buffer
            prev = [p for p in publisher.get(tile, []) if p[2] == k0]
            assert prev and prev[0][0] == cta - 1, (tile, cta, prev)
    for tile, pubs in publisher.items():
        for cta, _, k1 in pubs:                                      # every published partial is picked up
            assert any(c[1] == k1 and c[0] == cta + 1 for c in continuer.get(tile, [])), (tile, cta)
    if not force:
        chain = max((len(v) for v in publisher.values()), default=0)
        assert chain <= 2                                            # zgemm_auto: at most two hand-overs (three CTAs) per tile
        if total >= G_SMS:
            assert chain <= 1                                        # ranges of at least one tile: two CTAs per tile
        if streamk and total >= 2 * G_SMS:
            assert G_SMS < sk_tiles < 2 * G_SMS                      # one to two waves' worth, the rest in whole waves
