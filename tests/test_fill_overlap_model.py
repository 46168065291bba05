"""The two-stream schedule of a multi-batch fill (worldFill in csrc/mmgen.cu, mmgen_set_fill_overlap), restated as a happens-before graph.

CPU model, no GPU: the host enqueues, per batch b, on the FILL stream [wait evScan[b & 1] if b >= 2] gather(b) -> passes(b) ->
record evGather[b & 1], and on the MAIN stream wait evGather[b & 1] -> scan(b) -> record evScan[b & 1] -> deliver(b). A stream runs its
work in order; a wait refers to the most recent record of that event in host order (CUDA's rule). The test derives the ordering those
rules guarantee and checks that every pair of operations that touch the same buffer (one of them writing) is ordered, that nothing
waits in a cycle, and that the overlap the schedule exists for is really allowed: passes(b + 1) is NOT ordered after scan(b).
The GPU-side check of the same property is tests/test_gpu_parity.py::test_fill_overlap_is_result_neutral (identical chunk hashes).
"""
import itertools

import pytest


def build_schedule(n_batches, overlap=True, drop_scan_wait=False, drop_gather_wait=False):
    """Returns (ops, edges): ops[name] = (reads, writes) sets of buffer names; edges = happens-before pairs (a, b)."""
    ops, edges = {}, []
    last_on = {"fill": None, "main": None}
    last_record = {}                      # event name -> op after which it was last recorded (host order)

    def enqueue(stream, name, reads=(), writes=(), waits=()):
        ops[name] = (set(reads), set(writes))
        if last_on[stream] is not None:
            edges.append((last_on[stream], name))
        for ev in waits:
            if ev in last_record:
                edges.append((last_record[ev], name))
        last_on[stream] = name

    enqueue("main", "placements")         # everything before the fill; the fill stream forks from here (evSide[0])
    last_record["fork"] = "placements"
    first_fill_wait = ["fork"]
    for b in range(n_batches):
        s = b & 1
        sets = ["gathered%d" % s, "prep%d" % s]
        if overlap:
            waits = list(first_fill_wait)
            first_fill_wait = []
            if b >= 2 and not drop_scan_wait:
                waits.append("evScan%d" % s)
            enqueue("fill", "gather%d" % b, reads=["placement_lists"], writes=sets, waits=waits)
            enqueue("fill", "passes%d" % b, reads=["stage_products"], writes=["blocks%d" % b, "rock_queue", "lush_queue", "counters"])
            last_record["evGather%d" % s] = "passes%d" % b
            enqueue("main", "scan%d" % b, reads=sets, writes=["blocks%d" % b], waits=[] if drop_gather_wait else ["evGather%d" % s])
        else:
            # one stream: gather, passes, scan in sequence (mmgen_set_fill_overlap(0) with a single buffer set per parity)
            enqueue("main", "gather%d" % b, reads=["placement_lists"], writes=sets)
            enqueue("main", "passes%d" % b, reads=["stage_products"], writes=["blocks%d" % b, "rock_queue", "lush_queue", "counters"])
            enqueue("main", "scan%d" % b, reads=sets, writes=["blocks%d" % b])
        last_record["evScan%d" % s] = "scan%d" % b
        enqueue("main", "deliver%d" % b, reads=["blocks%d" % b])
    return ops, edges


def closure(ops, edges):
    names = list(ops)
    idx = {n: i for i, n in enumerate(names)}
    reach = [[False] * len(names) for _ in names]
    for a, b in edges:
        reach[idx[a]][idx[b]] = True
    for k in range(len(names)):
        rk = reach[k]
        for i in range(len(names)):
            if reach[i][k]:
                ri = reach[i]
                for j in range(len(names)):
                    if rk[j]:
                        ri[j] = True
    return names, idx, reach


def hazards(ops, edges):
    names, idx, reach = closure(ops, edges)
    assert not any(reach[i][i] for i in range(len(names))), "the waits form a cycle"
    bad = []
    for a, b in itertools.combinations(names, 2):
        ra, wa = ops[a]
        rb, wb = ops[b]
        if (wa & (rb | wb)) or (wb & ra):
            if not (reach[idx[a]][idx[b]] or reach[idx[b]][idx[a]]):
                bad.append((a, b))
    return bad, (names, idx, reach)


@pytest.mark.parametrize("n", [2, 3, 4, 7, 32])
def test_overlapped_schedule_orders_every_conflict(n):
    ops, edges = build_schedule(n)
    bad, (names, idx, reach) = hazards(ops, edges)
    assert not bad, bad
    for b in range(n - 1):
        # the point of the schedule: the next batch's passes may run while this batch's placements are scanned
        assert not reach[idx["scan%d" % b]][idx["passes%d" % (b + 1)]]
        assert not reach[idx["passes%d" % (b + 1)]][idx["scan%d" % b]]
    for b in range(n):
        assert reach[idx["passes%d" % b]][idx["scan%d" % b]] and reach[idx["scan%d" % b]][idx["deliver%d" % b]]
    # the fill stream never runs more than two batches ahead of the scan: the third gather reuses the first one's buffer set
    for b in range(2, n):
        assert reach[idx["scan%d" % (b - 2)]][idx["gather%d" % b]]
    # when the main stream is done the fill stream is done (the caller only synchronises the main stream)
    last = "deliver%d" % (n - 1)
    assert all(reach[idx[o]][idx[last]] for o in names if o != last)


def test_sequential_schedule_orders_every_conflict():
    ops, edges = build_schedule(5, overlap=False)
    bad, _ = hazards(ops, edges)
    assert not bad, bad


def test_the_model_sees_a_missing_wait():
    """Each of the two waits is necessary: without the scan wait the third gather overwrites lists the first scan may still read;
    without the gather wait a scan can start before its batch's terrain is there."""
    ops, edges = build_schedule(4, drop_scan_wait=True)
    bad, _ = hazards(ops, edges)
    assert ("scan0", "gather2") in bad or ("gather2", "scan0") in bad
    ops, edges = build_schedule(2, drop_gather_wait=True)
    bad, _ = hazards(ops, edges)
    assert any("passes0" in p and "scan0" in p for p in bad)
