"""CPU model of the mbarrier protocol of the halo conv kernel (srl_zoo_b200/csrc/conv_halo_tc.cu): eight producer warps, one MMA
issuer whose tcgen05 work completes asynchronously and in order, four epilogue warps, 21 phase-parity barriers (row_full[8] count 8,
row_free[8] count 1, tfull[2] count 1, tempty[2] count 4).  The agents below follow the kernel's wait / arrive order line by line
(producers :150-210, issuer :213-265, epilogue :285-320, plan :407-470) and are run under random interleavings with exact
`mbarrier.try_wait.parity` semantics; every shared-memory image row and TMEM accumulator carries a version, so that a read of a row
that is being rewritten, a rewrite of a row that an in-flight MMA still reads, an accumulator overwritten before it was drained, or a
deadlock fail the test.  (compute-sanitizer's synccheck reports "Missing init" for this kernel; profiles/r2_sanitizer.md shows
the report is reproduced by a minimal correct mbarrier program.  This model is the protocol argument behind that statement.)
A mutant that drops the row_free wait of the rows a warp does not write -- the hazard the kernel's comment at :159-163 names -- must be
caught by the same checks."""
import random

import pytest

MAXNR = 8


def make_plan(transposed, stride, pad, OH, OW):
    """port of make_plan (conv_halo_tc.cu:407-470) for 3x3 kernels: -> dict(HW, R, NR, ngroups, ncls, groups=[group of op o])"""
    s = stride
    out_s = s if transposed else 1
    ncls = out_s * out_s
    taps, maxow, maxoh = [], 0, 0
    for c in range(ncls):
        py, px = c // out_s, c % out_s
        maxoh = max(maxoh, (OH - py + out_s - 1) // out_s)
        maxow = max(maxow, (OW - px + out_s - 1) // out_s)
        for ky in range(3):
            for kx in range(3):
                if transposed:
                    ny, nx = py + pad - ky, px + pad - kx
                    if ny % s or nx % s:
                        continue
                    dy = ny // s if ny >= 0 else -((-ny) // s)
                    dx = nx // s if nx >= 0 else -((-nx) // s)
                else:
                    dy, dx = ky - pad, kx - pad
                taps.append((dy, dx, c))
    assert 0 < len(taps) <= 9
    mny, mxy = min(t[0] for t in taps), max(t[0] for t in taps)
    mnx, mxx = min(t[1] for t in taps), max(t[1] for t in taps)
    HW = maxow + (mxx - mnx)
    ngroups = mxy - mny + 1
    assert HW <= 112
    R = min(128 // HW, maxoh)
    NR = R + ngroups - 1
    if NR > MAXNR:
        R, NR = MAXNR - ngroups + 1, MAXNR
    assert R >= 1 and NR * 2 * HW <= 512
    groups = [g for g in range(ngroups) for t in taps if t[0] - mny == g]
    return dict(HW=HW, R=R, NR=NR, ngroups=ngroups, ncls=ncls, groups=groups)


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.phase, self.pending = self.phase + 1, self.count

    def done(self, parity):   # mbarrier.try_wait.parity: the phase of this parity has completed <=> the current phase has the other one
        return (self.phase & 1) != parity


class Hazard(AssertionError):
    pass


def simulate(plan, tiles, seed, mutant=False):
    rng = random.Random(seed)
    HW, R, NR, groups = plan["HW"], plan["R"], plan["NR"], plan["groups"]
    nops, ipr = len(groups), 2 * plan["HW"]
    nitems = NR * ipr
    row_full = [Bar(8) for _ in range(MAXNR)]
    row_free = [Bar(1) for _ in range(MAXNR)]
    tfull, tempty = [Bar(1), Bar(1)], [Bar(4), Bar(4)]
    # which producer warps store into row j (thread pidx owns items pidx and pidx + 256; a warp's step-k items are 256k + 32pw ..+31)
    owners = [set() for _ in range(NR)]
    for pw in range(8):
        for k in range(2):
            for i in range(256 * k + 32 * pw, min(256 * k + 32 * pw + 32, nitems)):
                owners[i // ipr].add(pw)
    assert all(owners)
    row_ver = [{w: -1 for w in owners[j]} for j in range(NR)]       # tile whose data warp w has stored in row j
    rows_of = [range(g, g + R) for g in groups]                     # rows whose content op o needs (output rows that are kept)
    last_reader = [max(o for o in range(nops) if j in rows_of[o]) for j in range(NR)]
    row_done = [-1] * NR                                            # last tile whose reads of row j have all COMPLETED
    acc_done = [-1, -1]                                             # last tile drained from accumulator buffer b by all 4 epilogue warps
    acc_drains = [0, 0]
    ops_completed = [0]
    pipe = []                                                       # in-order asynchronous tensor pipe: ("mma", it, o) | ("arrive", Bar)

    def producer(pw):
        for it in range(tiles):
            fph = it & 1
            arrived = 0

            def pass_rows(upto):
                nonlocal arrived
                while arrived < upto:
                    if not mutant:
                        yield row_free[arrived], fph ^ 1
                    row_full[arrived].arrive()
                    arrived += 1
            for k in range(2):
                lo_i = 256 * k + 32 * pw
                if lo_i >= nitems:
                    break
                lo_row, hi_row = lo_i // ipr, min((lo_i + 31) // ipr, NR - 1)
                yield from pass_rows(lo_row)
                for row in range(lo_row, hi_row + 1):
                    yield row_free[row], fph ^ 1
                    if it > 0 and row_done[row] != it - 1:
                        raise Hazard("row %d rewritten for tile %d while tile %d is still being read" % (row, it, it - 1))
                    row_ver[row][pw] = it
                    row_full[row].arrive()
                    arrived = row + 1
            yield from pass_rows(NR)

    def issuer():
        for it in range(tiles):
            buf, fph = it & 1, it & 1
            yield tempty[buf], ((it >> 1) & 1) ^ 1
            rows_ready = 0
            for o, g in enumerate(groups):
                while rows_ready < g + R:
                    yield row_full[rows_ready], fph
                    rows_ready += 1
                pipe.append(("mma", it, o))
                if o + 1 == nops or groups[o + 1] != g:
                    pipe.append(("arrive", row_free[g]))
                    if o + 1 == nops:
                        for j in range(g + 1, NR):
                            pipe.append(("arrive", row_free[j]))
                        pipe.append(("arrive", tfull[buf]))
                yield None   # a scheduling point between two issues

    def tensor_pipe():
        while True:
            while not pipe:
                yield "idle"
            item = pipe.pop(0)
            if item[0] == "arrive":
                item[1].arrive()
            else:
                _, it, o = item
                if it >= 2 and acc_done[it & 1] != it - 2:
                    raise Hazard("accumulator %d overwritten by tile %d before tile %d was drained" % (it & 1, it, it - 2))
                for j in rows_of[o]:
                    if any(v != it for v in row_ver[j].values()):
                        raise Hazard("tile %d op %d reads row %d holding %s" % (it, o, j, row_ver[j]))
                    if last_reader[j] == o:
                        row_done[j] = it
                ops_completed[0] += 1
            yield None

    def epilogue(w):
        for it in range(tiles):
            buf = it & 1
            yield tfull[buf], (it >> 1) & 1
            if ops_completed[0] < (it + 1) * nops:
                raise Hazard("epilogue reads tile %d before its MMAs completed" % it)
            yield None   # reading the accumulator takes time (an overwrite in the meantime is caught in tensor_pipe: acc_done)
            acc_drains[buf] += 1
            if acc_drains[buf] == 4:
                acc_drains[buf], acc_done[buf] = 0, it
            tempty[buf].arrive()

    agents = {("prod", w): producer(w) for w in range(8)}
    agents[("mma",)] = issuer()
    agents.update({("epi", w): epilogue(w) for w in range(4)})
    pipe_agent = tensor_pipe()
    waiting = {k: None for k in agents}     # the (barrier, parity) an agent is blocked on, or None = ready to run
    live = set(agents)
    for k, g in agents.items():
        try:
            waiting[k] = next(g)
        except StopIteration:
            live.discard(k)
    steps = 0
    while live:
        runnable = [k for k in live if waiting[k] is None or waiting[k][0].done(waiting[k][1])]
        if pipe:
            runnable.append("pipe")
        if not runnable:
            raise Hazard("deadlock: %s" % {k: (id(waiting[k][0]) % 1000, waiting[k][1]) for k in live})
        k = rng.choice(runnable)
        if k == "pipe":
            next(pipe_agent)
        else:
            try:
                waiting[k] = agents[k].send(None)
            except StopIteration:
                live.discard(k)
        steps += 1
        assert steps < 10_000_000
    while pipe:
        next(pipe_agent)
    assert ops_completed[0] == tiles * nops
    return steps


# the layer geometries the product routes through gconv64_halo_kernel (models/models.py:54,59,66-78)
GEOMETRIES = {"enc4 conv3x3 s1 56x56": (False, 1, 1, 56, 56), "dec0 convT 6->13": (True, 2, 0, 13, 13),
              "dec3 convT 13->27": (True, 2, 0, 27, 27), "dec6 convT 27->55": (True, 2, 0, 55, 55),
              "dec9 convT 55->111": (True, 2, 0, 111, 111), "enc8 dgrad 14->27 (conv3x3 s2 p1)": (True, 2, 1, 27, 27)}


def test_plan_port_matches_known_geometries():
    p = make_plan(*GEOMETRIES["dec0 convT 6->13"])
    assert (p["HW"], p["R"], p["NR"], p["ngroups"], p["ncls"], len(p["groups"])) == (8, 7, 8, 2, 4, 9)   # the geometry synccheck reports
    p = make_plan(*GEOMETRIES["enc4 conv3x3 s1 56x56"])
    assert (p["HW"], p["R"], p["NR"], p["ngroups"], p["ncls"], len(p["groups"])) == (58, 2, 4, 3, 1, 9)
    p = make_plan(*GEOMETRIES["dec9 convT 55->111"])
    assert (p["HW"], p["R"], p["NR"], p["ngroups"], p["ncls"]) == (57, 2, 3, 2, 4)


@pytest.mark.parametrize("name", list(GEOMETRIES))
def test_protocol_is_hazard_free_under_random_interleavings(name):
    plan = make_plan(*GEOMETRIES[name])
    for seed in range(40):
        simulate(plan, tiles=5, seed=seed)


def test_model_catches_the_hazard_the_kernel_comment_names():
    """without the row_free wait in pass_rows a fast warp's arrival completes the previous tile's phase of row_full in place of a
    slower warp's: the MMAs then read a row that is still being written (or the parity aliases into a deadlock)"""
    plan = make_plan(*GEOMETRIES["dec0 convT 6->13"])
    caught = 0
    for seed in range(40):
        try:
            simulate(plan, tiles=5, seed=seed, mutant=True)
        except Hazard:
            caught += 1
    assert caught >= 20, caught
