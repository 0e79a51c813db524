"""Length-bucketed batching: an opt-in that stops a mixed-length list from being padded to its
longest clip.

The reference pads every clip of a list call to the batch maximum (sylber/model/sylber.py:93-118); with 2-30 s clips
that is up to 15x wasted work, and because conv-0's GroupNorm runs over the padded axis the padding is
visible in the results.  Bucketing therefore changes results exactly as if the caller had split the list into the
same buckets and called the reference once per bucket - which is how the parity test states it."""
from __future__ import annotations


def plan_length_buckets(lengths, ratio=1.25, max_batch=64):
    """Partition clip indices into buckets whose longest / shortest length is <= `ratio`, at most `max_batch` each.

    Greedy on the lengths sorted in descending order (ties by index, so the plan is deterministic): a bucket opens at
    the longest unassigned clip and takes clips while they are at least `longest / ratio` long.  Returns a list of
    index lists; inside a bucket the original order is kept."""
    if ratio < 1.0:
        raise ValueError("ratio must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    buckets, cur, head = [], [], None
    for i in order:
        n = int(lengths[i])
        if cur and (len(cur) >= max_batch or n * ratio < head):
            buckets.append(sorted(cur))
            cur = []
        if not cur:
            head = n
        cur.append(i)
    if cur:
        buckets.append(sorted(cur))
    return buckets


def padded_work(lengths, buckets=None):
    """Samples processed (each clip padded to its bucket's maximum); `buckets=None` is the reference's single batch."""
    if buckets is None:
        return len(lengths) * max(lengths) if lengths else 0
    return sum(len(b) * max(lengths[i] for i in b) for b in buckets)


def sub_batch_bounds(n_rows, streams=3, max_batch=64, explicit=None):
    """(lo, hi) bounds of the sub-batches one padded batch of `n_rows` rows is run as (Segmenter._run_jobs).

    A batch of at most `max_batch` rows is split over up to `streams` sub-batches (none smaller than 8 rows on
    average) so that the copies of one overlap the kernels of the others; the FIRST sub-batch gets 2/3 of an even
    share because its host->device copy is the one nothing overlaps, while the last one's hidden states travel to
    the host during its own segmentation scan (Segmenter._run_jobs).  32 rows, 3 streams -> 8, 12, 12: measured 6.01-6.05 ms
    against 6.10 for 12, 12, 8 and 6.59 for 6, 10, 10, 6 (profiles/r03_e2e.md).  Longer lists go through in `max_batch`
    chunks.  `explicit` (a list of sizes that sums to n_rows) overrides the rule - used by tools/e2e_splits.py."""
    if explicit and sum(explicit) == n_rows:
        out, a = [], 0
        for k in explicit:
            out.append((a, a + k))
            a += k
        return out
    n_sub = max(1, min(streams, n_rows // 8)) if n_rows <= max_batch else 1
    out = []
    for lo in range(0, n_rows, max_batch):
        hi = min(lo + max_batch, n_rows)
        n = hi - lo
        sizes = [int(round(n / (n_sub - 1 / 3)))] * (n_sub - 1) if n_sub > 1 else []
        sizes.insert(0, n - sum(sizes))
        if min(sizes) <= 0:
            sizes = [n]
        a = lo
        for k in sizes:
            out.append((a, a + k))
            a += k
    return out
