"""Host model of the tile-ticket arithmetic of k_classify<early> (arpeggio_b200/csrc/arp_pairs.cu): counter c serves
the tickets and the blocks congruent to c modulo nc = min(ARP_CLS_COUNTERS, gridDim); a warp's first ticket is
static, later ones are (class_warps + t) * nc + c for the t-th fetch from counter c.  Every tile index must be handed
out exactly once for every grid size (a grid smaller than the number of counters once left whole classes unserved)."""
import pytest

CLS_WARPS = 8
ARP_CLS_COUNTERS = 16


def handed_out(grid, n_tiles):
    nc = min(ARP_CLS_COUNTERS, grid)
    seen = []
    for cls in range(nc):
        blocks = [b for b in range(grid) if b % nc == cls]
        class_warps = ((grid - cls + nc - 1) // nc) * CLS_WARPS
        assert class_warps == len(blocks) * CLS_WARPS
        first = [((b // nc) * CLS_WARPS + w) * nc + cls for b in blocks for w in range(CLS_WARPS)]
        seen += [t for t in first if t < n_tiles]
        t = 0
        while (class_warps + t) * nc + cls < n_tiles:           # the counter keeps counting until the tiles run out
            seen.append((class_warps + t) * nc + cls)
            t += 1
    return seen


@pytest.mark.parametrize('grid', [1, 2, 3, 7, 15, 16, 17, 31, 100, 521, 592])
@pytest.mark.parametrize('n_tiles', [0, 1, 5, 127, 128, 129, 4736, 20555])
def test_every_tile_exactly_once(grid, n_tiles):
    seen = handed_out(grid, n_tiles)
    assert sorted(seen) == list(range(n_tiles))
