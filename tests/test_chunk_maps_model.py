"""Host model of the chunked sampler's arithmetic (csrc/ransac.cu, section 1b) on the real cv::RNG stream.

The serial getSubset walks the draw stream as a chain p -> p + consumed(p), where consumed(p) — the draws an attempt
starting at position p uses up until it holds four distinct indices — depends on the draws alone.  The CUDA kernels cut
the stream into chunks, compute per chunk the map "entry offset -> exit offset into the next chunk" for every entry offset
below CH_K, and compose the maps in order.  This test restates that in numpy and checks, for several set sizes, that
  * the composed entries are the positions the serial chain really has at the chunk boundaries,
  * every exit offset stays below CH_K (the kernels hand the set to the serial sampler otherwise),
  * chains started at different offsets do NOT generally meet inside a window when n is large (why a first version that
    waited for convergence always fell back).
No GPU, no oracle: the RNG is three lines."""
import numpy as np
import pytest

CH_K = 64
CHUNK = 4096            # the kernels use 16 384; the arithmetic does not depend on the size


def rng_stream(count):
    """Raw outputs of cv::RNG seeded with (uint64)-1: state = lo32 * 4164903690 + hi32."""
    out = np.empty(count, np.uint32)
    state = 0xFFFFFFFFFFFFFFFF
    for i in range(count):
        state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & 0xFFFFFFFFFFFFFFFF
        out[i] = state & 0xFFFFFFFF
    return out


def consumed_all(draws):
    """consumed(p) for every p that has enough look-ahead (0 = would overrun)."""
    n = len(draws)
    cons = np.zeros(n, np.int32)
    for p in range(n - 80):
        idx = []
        q = p
        while len(idx) < 4:
            v = draws[q]; q += 1
            if v not in idx:
                idx.append(v)
        cons[p] = q - p
    return cons


@pytest.fixture(scope="module")
def raw():
    return rng_stream(6 * CHUNK + 200)


@pytest.mark.parametrize("n", [5, 6, 32, 600, 8192])
def test_entry_maps_compose_to_the_serial_chain(raw, n):
    draws = (raw % np.uint32(n)).astype(np.int64)
    cons = consumed_all(list(draws))
    n_chunks = 5
    # the serial chain from position 0
    chain = []
    p = 0
    while p < n_chunks * CHUNK:
        chain.append(p); p += cons[p]
    chain = np.array(chain + [p])
    true_entry = [int(chain[np.searchsorted(chain, c * CHUNK)]) - c * CHUNK for c in range(n_chunks + 1)]
    assert true_entry[0] == 0
    # per-chunk maps: entry offset e -> first chain position at or beyond the chunk's end, as an offset
    entry = 0
    for c in range(n_chunks):
        begin, end = c * CHUNK, (c + 1) * CHUNK
        emap = np.full(CH_K, -1)
        for e in range(CH_K):
            p = begin + e
            while p < end:
                p += cons[p]
            if p - end < CH_K:
                emap[e] = p - end
        assert entry == true_entry[c], (n, c)
        assert emap[entry] >= 0, (n, c)              # the kernels' "unsafe" case must not occur on this stream
        entry = int(emap[entry])
    assert entry == true_entry[n_chunks]
    assert max(true_entry) < CH_K


def test_chains_of_different_phase_do_not_meet_for_large_n(raw):
    """n = 8192: an attempt consumes exactly four draws unless two of them collide (probability ~6 / 8192), so chains
    started at offsets 0..3 stay apart for thousands of draws; with n = 32 they merge within a few attempts."""
    def distinct_after(n, length):
        draws = list((raw[:length + 200] % np.uint32(n)).astype(np.int64))
        cons = consumed_all(draws)
        ends = set()
        for e in range(4):
            p = e
            while p < length:
                p += cons[p]
            ends.add(p)
        return len(ends)
    assert distinct_after(8192, 2048) > 1
    assert distinct_after(32, 2048) == 1
