"""Model of one CTA's weight ring in k_decode_persistent: the producer walks the tiles of several matrices back to back
(pd_produce_matrix), eight consumer warps pop "their" tiles (pd_consume_matrix: seq = seq_base + s * nw + warp, release
counter rel[stage] == lap, full-barrier parity lap & 1) with a CTA barrier between matrices.  Random interleavings; checks that
every warp finds exactly the tile it expects in the stage it looks at, and that nothing deadlocks.
usage: python tools/sim/decode_persistent_ring.py [runs]"""
import random
import sys

W_CONS = 8


def tiles_of(g0, g1, nseg):
    """Producer order for one matrix: (group, segment) per seq."""
    out = []
    gb = g0
    while gb < g1:
        nw = min(W_CONS, g1 - gb)
        for s in range(nseg):
            for w in range(nw):
                out.append((gb + w, s))
        gb += W_CONS
    return out


def simulate(S, mats, rng):
    full_phase = [0] * S      # full barrier phase bit per stage
    full_armed = [False] * S
    empty_phase = [0] * S
    rel = [0] * S
    stage = [None] * S        # (matrix, group, segment) currently in the stage
    arrived = {"n": 0, "gen": 0}

    def producer():
        seq = 0
        for m, (g0, g1, nseg) in enumerate(mats):
            for (g, s) in tiles_of(g0, g1, nseg):
                st, ph = seq % S, (seq // S) & 1
                while not (empty_phase[st] != (ph ^ 1)):  # mbar_wait(empty, ph ^ 1)
                    yield
                yield  # copy in flight
                stage[st] = (m, g, s)
                full_phase[st] ^= 1  # expect_tx + complete_tx: phase completes
                seq += 1
                yield

    def consumer(w):
        seq_base = 0
        for m, (g0, g1, nseg) in enumerate(mats):
            gb = g0
            while gb < g1:
                nw = min(W_CONS, g1 - gb)
                if w < nw:
                    for s in range(nseg):
                        seq = seq_base + s * nw + w
                        st, lap = seq % S, seq // S
                        while rel[st] != lap:
                            yield
                        while not (full_phase[st] != (lap & 1)):  # mbar_wait(full, lap & 1)
                            yield
                        assert stage[st] == (m, gb + w, s), f"warp {w} expects {(m, gb + w, s)} in stage {st}, finds {stage[st]}"
                        yield  # dot products
                        rel[st] = lap + 1
                        empty_phase[st] ^= 1  # mbar_arrive(empty)
                        yield
                seq_base += nseg * nw
                gb += W_CONS
            # pd_arrive: CTA barrier over the consumer warps before the next matrix
            gen = arrived["gen"]
            arrived["n"] += 1
            if arrived["n"] == W_CONS:
                arrived["n"] = 0
                arrived["gen"] += 1
            while arrived["gen"] == gen:
                yield

    procs = {"p": producer(), **{w: consumer(w) for w in range(W_CONS)}}
    stalled = 0
    while procs:
        k = rng.choice(list(procs))
        snap = (tuple(rel), tuple(full_phase), tuple(empty_phase), arrived["gen"], arrived["n"])
        try:
            next(procs[k])
        except StopIteration:
            del procs[k]
            stalled = 0
            continue
        stalled = 0 if (tuple(rel), tuple(full_phase), tuple(empty_phase), arrived["gen"], arrived["n"]) != snap else stalled + 1
        assert stalled < 100000, f"deadlock: {list(procs)} (S={S}, mats={mats})"


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = random.Random(11)
    for _ in range(runs):
        S = rng.choice([3, 4, 7, 10, 20])
        mats = []
        for _m in range(rng.choice([1, 2, 4, 6])):
            g0 = rng.randrange(0, 5)
            mats.append((g0, g0 + rng.randrange(0, 30), rng.choice([1, 2, 7])))
        simulate(S, mats, rng)
    print(f"{runs} runs ok: every warp finds its tile, no deadlock, ring indices continuous across matrices")
    return 0


if __name__ == "__main__":
    sys.exit(main())
