"""Discrete simulation of k_join_pipe's mbarrier protocol (csrc/k_join_pipe.cu) -- CPU-only design check.

Models the PTX mbarrier semantics the kernel relies on (phase completes when the pending arrival count AND the pending
transaction bytes reach zero; try_wait.parity(p) succeeds iff the phase with parity p has completed, i.e. the barrier's
current phase parity differs from p) and runs the producer warp, the four consumer warps and the asynchronous copies
under a random scheduler, for many (batches, chunks, stages, empty-batch patterns).  Checks: no deadlock, no barrier ever
receives more arrivals than it expects in a phase, every consumer reads exactly the (batch, chunk) the producer put into
the stage / descriptor slot it reads, and no slot is overwritten before all consumers released it.

    python scripts/pipe_protocol_sim.py [trials]
"""
import random
import sys

CW, NM, MDIST = 4, 4, 2


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        assert self.pending > 0, "more arrivals than the barrier expects in this phase"
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, n):
        self.tx -= n
        self._maybe_complete()

    def test(self, parity):
        return (self.phase & 1) != parity


def run(nmy, nchunks, NST, empties, rng):
    full = [MBar(1) for _ in range(NST)]; empty = [MBar(CW) for _ in range(NST)]
    mfull = [MBar(1) for _ in range(NM)]; mempty = [MBar(CW) for _ in range(NM)]
    stage_content = [None] * NST; stage_readers = [0] * NST
    meta_content = [None] * NM; meta_readers = [0] * NM
    inflight = []                                    # pending async copies: (kind, slot, payload, bytes)

    def producer():
        def fetch(i):
            q, n = i % NM, i // NM
            if n >= 1:
                while not mempty[q].test((n - 1) & 1):
                    yield
            mfull[q].arrive(expect_tx=1)
            inflight.append(("meta", q, i, 1))
        for i in range(min(MDIST, nmy)):
            yield from fetch(i)
        it = 0
        for i in range(nmy):
            if i + MDIST < nmy:
                yield from fetch(i + MDIST)
            q = i % NM
            while not mfull[q].test((i // NM) & 1):
                yield
            assert meta_content[q] == i, f"producer read descriptor {meta_content[q]} expecting {i}"
            if empties[i]:
                continue
            for c in range(nchunks):
                s, n = it % NST, it // NST
                if n >= 1:
                    while not empty[s].test((n - 1) & 1):
                        yield
                full[s].arrive(expect_tx=3)
                yield
                for part in range(3):                # several bulk copies per stage, completing in any order
                    inflight.append(("stage", s, (i, c), 1))
                    yield
                it += 1

    def consumer(w):
        it = 0
        for i in range(nmy):
            q = i % NM
            while not mfull[q].test((i // NM) & 1):
                yield
            assert meta_content[q] == i, f"consumer {w} read descriptor {meta_content[q]} expecting {i}"
            meta_readers[q] += 1
            if not empties[i]:
                for c in range(nchunks):
                    s = it % NST
                    while not full[s].test((it // NST) & 1):
                        yield
                    assert stage_content[s] == (i, c), f"consumer {w} read stage {stage_content[s]} expecting {(i, c)}"
                    stage_readers[s] += 1
                    yield
                    empty[s].arrive()
                    it += 1
            yield
            mempty[q].arrive()

    actors = [producer()] + [consumer(w) for w in range(CW)]
    alive = [True] * len(actors)
    idle_rounds = 0
    while any(alive):
        progressed = False
        # asynchronous copies land in arbitrary order
        if inflight and rng.random() < 0.6:
            kind, slot, payload, nbytes = inflight.pop(rng.randrange(len(inflight)))
            if kind == "meta":
                assert meta_content[slot] is None or meta_readers[slot] == CW, "descriptor overwritten before release"
                meta_content[slot] = payload; meta_readers[slot] = 0
                mfull[slot].complete_tx(nbytes)
            else:
                if stage_content[slot] != payload:
                    assert stage_content[slot] is None or stage_readers[slot] == CW, "stage overwritten before release"
                    stage_content[slot] = payload; stage_readers[slot] = 0
                full[slot].complete_tx(nbytes)
            progressed = True
        a = rng.randrange(len(actors))
        if alive[a]:
            before = (tuple(b.phase for b in full + empty + mfull + mempty), len(inflight))
            try:
                next(actors[a])
            except StopIteration:
                alive[a] = False
                progressed = True
            after = (tuple(b.phase for b in full + empty + mfull + mempty), len(inflight))
            progressed |= before != after
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 20000, f"deadlock: nmy={nmy} nchunks={nchunks} NST={NST}"
    assert not inflight


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    rng = random.Random(1)
    for t in range(trials):
        nmy = rng.choice([0, 1, 2, 3, 4, 5, 7, 9, 17])
        nchunks = rng.choice([1, 2, 3, 4, 8])
        NST = rng.choice([2, 3, 4, 6])
        empties = [rng.random() < 0.15 for _ in range(nmy)]
        run(nmy, nchunks, NST, empties, rng)
    print(f"pipe protocol: {trials} randomised schedules OK")


if __name__ == "__main__":
    main()
