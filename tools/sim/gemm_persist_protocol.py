"""Discrete-event model of the synchronisation protocol of k_gemm_f16_2cta_persist (csrc/prefill_gemm.cuh): the operand ring
(full / empty mbarriers, cta_group::2 loads completing on the leader, multicast commits) and the double-buffered TMEM
accumulator (tmem_full / tmem_empty) of one CTA pair, with the exact parity expressions of the kernel, run under random
interleavings.  Checks: no deadlock, the MMA thread never touches a stage before both CTAs' bytes landed, a producer never
overwrites a stage the MMA has not finished, an epilogue never reads an accumulator before its tile is complete, the MMA
never overwrites an accumulator before BOTH epilogues drained its previous tile.
usage: python tools/sim/gemm_persist_protocol.py [runs]"""
import random
import sys


class MBar:
    """mbarrier: phase bit, pending arrivals, transaction count (may go negative transiently, as in hardware)."""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        self.tx += expect_tx
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals in one phase"
        self._maybe_flip()

    def complete_tx(self, n):
        self.tx -= n
        self._maybe_flip()

    def test(self, parity):  # try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


def simulate(S, nk, tiles, rng):
    A_B = 10  # bytes one CTA loads per stage (symbolic)
    full = [MBar(1) for _ in range(S)]                       # leader's full barriers
    empty = [[MBar(1) for _ in range(S)] for _ in range(2)]  # per CTA
    tfull = [[MBar(1) for _ in range(2)] for _ in range(2)]  # per CTA
    tempty = [MBar(8) for _ in range(2)]                     # leader's
    stage_data = [[None] * S for _ in range(2)]              # (tile, kb) whose bytes sit in CTA r's stage
    stage_busy = [[False] * S for _ in range(2)]             # loaded and not yet released by a commit
    acc_tile = [None, None]                                  # tile accumulated (complete) in TMEM buffer b
    acc_drained = [[True, True], [True, True]]               # [buf][cta]: epilogue finished reading
    log = {"mma": 0, "epi": [0, 0]}

    def producer(r):
        it = 0
        for t in range(tiles):
            for kb in range(nk):
                st = it % S
                while not empty[r][st].test(((it // S) & 1) ^ 1):
                    yield
                assert not stage_busy[r][st], f"producer {r} overwrites a live stage"
                if r == 0:
                    full[st].arrive(expect_tx=2 * A_B)
                yield  # TMA in flight
                stage_data[r][st] = (t, kb)
                stage_busy[r][st] = True
                full[st].complete_tx(A_B)
                it += 1
                yield

    def mma():
        it = 0
        for i in range(tiles):
            buf = i & 1
            while not tempty[buf].test(((i >> 1) & 1) ^ 1):
                yield
            assert all(acc_drained[buf]), "MMA overwrites an accumulator an epilogue is still reading"
            acc_tile[buf] = None
            for kb in range(nk):
                st = it % S
                while not full[st].test((it // S) & 1):
                    yield
                for r in range(2):
                    assert stage_data[r][st] == (i, kb), f"MMA reads stage {st} holding {stage_data[r][st]}, wants {(i, kb)}"
                yield  # the MMAs execute
                for r in range(2):  # commit, multicast to both CTAs
                    stage_busy[r][st] = False
                    empty[r][st].arrive()
                it += 1
                log["mma"] += 1
                yield
            acc_tile[buf] = i
            acc_drained[buf] = [False, False]
            for r in range(2):
                tfull[r][buf].arrive()
            yield

    def epilogue(r):
        for i in range(tiles):
            buf = i & 1
            while not tfull[r][buf].test((i >> 1) & 1):
                yield
            assert acc_tile[buf] == i, f"epilogue {r} reads buffer {buf} holding tile {acc_tile[buf]}, wants {i}"
            yield  # tcgen05.ld of the whole tile
            assert acc_tile[buf] == i, "accumulator overwritten while being read"
            acc_drained[buf][r] = True
            for _ in range(4):  # one remote arrive per epilogue warp
                tempty[buf].arrive()
            log["epi"][r] += 1
            yield  # stores

    procs = {"p0": producer(0), "p1": producer(1), "mma": mma(), "e0": epilogue(0), "e1": epilogue(1)}
    stalled = 0
    while procs:
        name = rng.choice(list(procs))
        before = (log["mma"], tuple(log["epi"]), tuple(b.phase for b in full), tuple(b.phase for b in tempty))
        try:
            next(procs[name])
        except StopIteration:
            del procs[name]
            stalled = 0
            continue
        after = (log["mma"], tuple(log["epi"]), tuple(b.phase for b in full), tuple(b.phase for b in tempty))
        stalled = 0 if after != before else stalled + 1
        assert stalled < 20000, f"deadlock: {list(procs)} stuck (S={S}, nk={nk}, tiles={tiles})"
    assert log["mma"] == tiles * nk and log["epi"] == [tiles, tiles]


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = random.Random(7)
    for _ in range(runs):
        simulate(rng.choice([2, 3, 5, 6]), rng.choice([1, 2, 3, 7, 16]), rng.choice([1, 2, 3, 4, 5, 9]), rng)
    print(f"{runs} random interleavings ok: no deadlock, no stage or accumulator hazard")
    return 0


if __name__ == "__main__":
    sys.exit(main())
