"""Host-side model of the row-slab peer-memory protocol (eq_b200/csrc/slab.cu: k_halo_push, k_peer_allreduce).

The kernels synchronise only through what they write into each other's memory: staging slots that carry their own
exchange number q, double-buffered by the parity of q, and all-reduce slots tagged q ^ bits(value) ^ salt, also by parity.
The argument in the source -- "parity q is free again when exchange q comes round, because a rank that starts q has
completed q-1, so it has received its neighbours' rows of q-1, which they sent after finishing q-2" -- is checked here by
brute force: N ranks run the same sequence of exchanges and all-reduces as independent state machines, a seeded scheduler
picks which rank advances (so ranks drift as far apart as the protocol allows), remote stores land after random delays and
out of order, and every read asserts that it sees the value of ITS exchange and that no slot is overwritten before its
reader is done with it.  No GPU, no library: this pins the reasoning, the CUDA tests (tests/test_slab_gpu.py) pin the code."""
import random

import pytest

SALT = 0x9E3779B97F4A7C15
MASK = (1 << 64) - 1


class Rank:
    def __init__(self, r, world):
        self.r, self.world = r, world
        # staging[side][parity] = list of (q, payload) slots; side 0 = from the lower neighbour, 1 = from the upper one
        self.staging = [[None, None], [None, None]]
        self.ar_slots = [[None, None] for _ in range(world)]   # [sender][parity] = (bits, tag)
        self.pc = 0            # index into the op sequence
        self.phase = 0         # 0: not sent yet, 1: sent, waiting
        self.xq = 0            # exchanges issued
        self.aq = 0            # all-reduces issued
        self.reading = None    # (kind, parity) while a kernel of this rank is between its first and last read


def run(world, ops, seed, max_delay):
    rng = random.Random(seed)
    ranks = [Rank(r, world) for r in range(world)]
    in_flight = []   # (deliver_at, kind, dst, key, value)
    clock = 0
    results = [[] for _ in range(world)]

    def post(kind, dst, key, value):
        in_flight.append((clock + rng.randint(0, max_delay), kind, dst, key, value))

    def deliver():
        nonlocal in_flight
        rest = []
        rng.shuffle(in_flight)   # posted stores of different kernels may land in any order
        for item in in_flight:
            at, kind, dst, key, value = item
            if at > clock:
                rest.append(item)
                continue
            R = ranks[dst]
            if kind == "halo":
                side, par = key
                # the hazard the parity argument excludes: overwriting rows the receiver has not consumed yet
                old = R.staging[side][par]
                if old is not None:
                    assert old[0] == value[0] - 2, ("staging overwritten out of turn", dst, side, old, value)
                    assert R.xq > old[0] or (R.xq == old[0] and R.phase == 0 and R_op_done(R, "x", old[0])), \
                        ("staging overwritten before its reader finished", dst, side, old, value, R.xq)
                R.staging[side][par] = value
            else:
                sender, par = key
                old = R.ar_slots[sender][par]
                if old is not None:
                    q_old = (old[1] ^ old[0] ^ SALT) & MASK
                    q_new = (value[1] ^ value[0] ^ SALT) & MASK
                    assert q_old == q_new - 2, ("all-reduce slot overwritten out of turn", dst, sender, q_old, q_new)
                    assert R_op_done(R, "a", q_old), ("all-reduce slot overwritten before it was read", dst, sender, q_old)
                R.ar_slots[sender][par] = value
        in_flight = rest

    done_ops = [dict() for _ in range(world)]

    def R_op_done(R, kind, q):
        return done_ops[R.r].get((kind, q), False)

    def step(R):
        if R.pc >= len(ops):
            return False
        kind = ops[R.pc]
        if kind == "x":
            if R.phase == 0:
                R.xq += 1
                q, par = R.xq, R.xq & 1
                if R.r > 0:
                    post("halo", R.r - 1, (1, par), (q, ("rows", R.r, q)))   # I am the upper neighbour of rank-1
                if R.r + 1 < R.world:
                    post("halo", R.r + 1, (0, par), (q, ("rows", R.r, q)))
                R.phase = 1
                return True
            q, par = R.xq, R.xq & 1
            need = [(0, R.r - 1)] if R.r > 0 else []
            need += [(1, R.r + 1)] if R.r + 1 < R.world else []
            for side, nb in need:
                slot = R.staging[side][par]
                if slot is None or slot[0] != q:
                    assert slot is None or slot[0] < q, ("a later exchange landed in the parity being waited for", R.r, slot, q)
                    return False   # keep polling
            for side, nb in need:
                assert R.staging[side][par] == (q, ("rows", nb, q))
            results[R.r].append(("x", q))
            done_ops[R.r][("x", q)] = True
            R.phase = 0
            R.pc += 1
            return True
        # all-reduce
        if R.phase == 0:
            R.aq += 1
            q, par = R.aq, R.aq & 1
            bits = (R.r * 1000003 + q * 7919) & MASK          # this rank's partial of all-reduce q
            for dst in range(R.world):
                post("ar", dst, (R.r, par), (bits, (q ^ bits ^ SALT) & MASK))
            R.phase = 1
            return True
        q, par = R.aq, R.aq & 1
        vals = []
        for sender in range(R.world):
            slot = R.ar_slots[sender][par]
            if slot is None or ((slot[1] ^ slot[0] ^ SALT) & MASK) != q:
                return False
            vals.append(slot[0])
        assert vals == [(s * 1000003 + q * 7919) & MASK for s in range(R.world)], ("all-reduce read a stale partial", R.r, q)
        results[R.r].append(("a", sum(vals)))
        done_ops[R.r][("a", q)] = True
        R.phase = 0
        R.pc += 1
        return True

    idle = 0
    while any(R.pc < len(ops) for R in ranks):
        clock += 1
        deliver()
        # a biased scheduler: one rank is favoured for a while, so it runs ahead as far as the protocol lets it
        fav = rng.randrange(world)
        order = [fav] * 4 + [rng.randrange(world) for _ in range(2)]
        progressed = False
        for r in order:
            progressed = step(ranks[r]) or progressed
        idle = 0 if progressed or in_flight else idle + 1
        assert idle < 10000, "protocol deadlocked"
    return results


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("max_delay", [0, 3, 40])
def test_protocol_never_reads_stale_or_overwrites_unread(world, max_delay):
    # one PCG iteration of the slab path: 12 exchanges and 3 all-reduces, interleaved as enqueue_fused_iteration issues them
    iteration = ["x"] * 11 + ["a", "x", "a", "a"]
    ops = (["a", "x"] + iteration * 3) * 2
    for seed in range(25):
        res = run(world, ops, seed, max_delay)
        for r in range(1, world):
            assert res[r] == res[0]   # every rank saw every exchange, and the same all-reduce sums in the same order


def test_ranks_drift_by_at_most_one_exchange():
    """The bound the parity argument rests on: a rank can be at most one exchange ahead of a neighbour."""
    world, ops = 4, ["x"] * 40
    rng = random.Random(5)
    ranks = [0] * world   # completed exchanges
    sent = [0] * world
    worst = 0
    for _ in range(20000):
        r = rng.choice([0, 0, 0, 1, 2, 3])   # rank 0 is pushed hard
        if sent[r] == ranks[r]:
            if sent[r] < len(ops):
                sent[r] += 1                 # phase 1 never waits
        else:
            nbs = [n for n in (r - 1, r + 1) if 0 <= n < world]
            if all(sent[n] >= sent[r] for n in nbs):
                ranks[r] = sent[r]
        for a in range(world - 1):
            worst = max(worst, abs(sent[a] - sent[a + 1]))
    assert worst <= 1
