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
