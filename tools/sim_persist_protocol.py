"""Discrete-event simulation of the barrier protocol of csrc/gemm_persist.inc (the persistent GEMM was written without a
GPU at hand): the producer / MMA / two epilogue-group loops are transcribed with the SAME stage, phase and parity
expressions as the kernel and run as coroutines against a model of mbarrier semantics (pending arrivals, transaction
bytes, phase bit; try_wait.parity(p) succeeds when the phase with parity p has completed) with random latencies.
Checked: no deadlock, every tile is accumulated and drained exactly once and in order, a ring stage is never refilled
before its MMAs retired, an accumulator is never overwritten before the epilogue group has read it, and a group's
staging row is not reused before its bulk store has been read.

    python tools/sim_persist_protocol.py          # a few thousand random configurations
"""
import random


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in one phase"
        self.pending -= 1
        self._maybe_flip()

    def expect_tx_arrive(self, n):
        self.tx += n
        self.arrive()

    def complete_tx(self, n):
        self.tx -= n
        self._maybe_flip()

    def done(self, parity):  # try_wait.parity
        return self.phase != parity


class Sim:
    def __init__(self, rng, stages, total_kb, n_local_tiles, with_residual):
        self.rng, self.S, self.KB, self.T, self.res = rng, stages, total_kb, n_local_tiles, with_residual
        self.full = [MBar(1) for _ in range(stages)]
        self.empty = [MBar(1) for _ in range(stages)]
        self.acc_full = [MBar(1), MBar(1)]
        self.acc_empty = [MBar(4), MBar(4)]
        self.res_bar = [MBar(1) for _ in range(8)]
        self.events = []            # (time, callback)
        self.now = 0
        self.stage_state = ["free"] * stages       # free -> loading -> loaded -> consumed(free)
        self.acc_state = ["free", "free"]          # free -> accumulating(tile) -> full(tile) -> free
        self.acc_tile = [None, None]
        self.stage_row_busy = [[False] * 4, [False] * 4]   # per group, per warp: bulk store still reading the staging row
        self.mma_tiles, self.drained = [], []

    def later(self, dt, fn):
        self.events.append((self.now + dt, self.rng.random(), fn))

    # ---- coroutines: yield ("wait", bar, parity) | ("sleep", n)
    def producer(self):
        it = 0
        for _tile in range(self.T):
            for _kb in range(self.KB):
                stage, phase = it % self.S, (it // self.S) & 1
                yield ("wait", self.empty[stage], phase ^ 1)
                assert self.stage_state[stage] == "free", "ring stage refilled before its MMAs retired"
                self.stage_state[stage] = "loading"
                self.full[stage].expect_tx_arrive(100)

                def landed(stage=stage):
                    self.stage_state[stage] = "loaded"
                    self.full[stage].complete_tx(100)
                self.later(self.rng.randint(1, 40), landed)
                yield ("sleep", 1)
                it += 1

    def mma(self):
        it = 0
        for i in range(self.T):
            acc, use = i & 1, i >> 1
            yield ("wait", self.acc_empty[acc], (use & 1) ^ 1)
            assert self.acc_state[acc] == "free", "accumulator overwritten before the epilogue group read it"
            self.acc_state[acc], self.acc_tile[acc] = "accumulating", i
            for _kb in range(self.KB):
                stage, phase = it % self.S, (it // self.S) & 1
                yield ("wait", self.full[stage], phase)
                assert self.stage_state[stage] == "loaded"

                def retired(stage=stage):  # tcgen05.commit -> empty barrier when these MMAs retire
                    self.stage_state[stage] = "free"
                    self.empty[stage].arrive()
                self.later(self.rng.randint(1, 12), retired)
                yield ("sleep", 1)
                it += 1

            def acc_done(acc=acc, i=i):
                assert self.acc_state[acc] == "accumulating" and self.acc_tile[acc] == i
                self.acc_state[acc] = "full"
                self.acc_full[acc].arrive()
            self.later(self.rng.randint(12, 20), acc_done)  # after every MMA of the tile (> the per-stage retire latency)
            self.mma_tiles.append(i)

    def epilogue_warp(self, eg, w):
        res_uses = 0
        for i in range(self.T):
            if (i & 1) != eg:
                continue
            use = i >> 1
            # bulk_wait_read0(): my previous store must have been read
            while self.stage_row_busy[eg][w]:
                yield ("sleep", 1)
            yield ("sleep", self.rng.randint(1, 10))  # tables + group barriers (not modelled: plain bar.sync)
            bar = self.res_bar[eg * 4 + w]
            if self.res:
                bar.expect_tx_arrive(64)
                self.later(self.rng.randint(1, 60), lambda bar=bar: bar.complete_tx(64))
            yield ("wait", self.acc_full[eg], use & 1)
            assert self.acc_state[eg] == "full" and self.acc_tile[eg] == i, "epilogue read the wrong accumulator contents"
            if self.res:
                yield ("wait", bar, res_uses & 1)
                res_uses += 1
            yield ("sleep", self.rng.randint(1, 30))  # TMEM reads
            self.acc_reads[eg] += 1
            if self.acc_reads[eg] == 4:  # the 4th warp's arrival completes the phase: buffer free again
                self.acc_reads[eg] = 0
                self.acc_state[eg] = "free"
                self.drained.append(i)
            self.acc_empty[eg].arrive()
            yield ("sleep", self.rng.randint(1, 30))  # math + staging writes
            self.stage_row_busy[eg][w] = True

            def store_read(eg=eg, w=w):
                self.stage_row_busy[eg][w] = False
            self.later(self.rng.randint(1, 80), store_read)

    def run(self):
        self.acc_reads = [0, 0]
        threads = [self.producer(), self.mma()] + [self.epilogue_warp(eg, w) for eg in range(2) for w in range(4)]
        state = [None] * len(threads)   # pending request per thread
        alive = [True] * len(threads)
        wake = [0] * len(threads)
        for step in range(2_000_000):
            progressed = False
            for k, th in enumerate(threads):
                if not alive[k]:
                    continue
                req = state[k]
                if req is not None:
                    if req[0] == "wait" and not req[1].done(req[2]):
                        continue
                    if req[0] == "sleep" and self.now < wake[k]:
                        continue
                try:
                    state[k] = next(th)
                    if state[k][0] == "sleep":
                        wake[k] = self.now + state[k][1]
                    progressed = True
                except StopIteration:
                    alive[k] = False
                    progressed = True
            if not any(alive) and not self.events:
                break
            self.now += 1
            due = sorted([e for e in self.events if e[0] <= self.now])
            self.events = [e for e in self.events if e[0] > self.now]
            for _, _, fn in due:
                fn()
            if not progressed and not due and not self.events and all(
                    state[k] is not None and state[k][0] == "wait" and not state[k][1].done(state[k][2])
                    for k in range(len(threads)) if alive[k]):
                raise AssertionError(f"deadlock at t={self.now}: "
                                     f"{[(k, state[k][2]) for k in range(len(threads)) if alive[k]]}")
        else:
            raise AssertionError("simulation did not finish")
        assert self.mma_tiles == list(range(self.T)), self.mma_tiles
        assert sorted(self.drained) == list(range(self.T)), self.drained
        assert all(s == "free" for s in self.stage_state) and self.acc_state == ["free", "free"]


def main(n=3000, seed=0):
    rng = random.Random(seed)
    for _ in range(n):
        Sim(rng, stages=rng.randint(2, 6), total_kb=rng.randint(1, 24), n_local_tiles=rng.randint(1, 9),
            with_residual=rng.random() < 0.5).run()
    return n


if __name__ == "__main__":
    print("persistent-GEMM barrier protocol: %d random configurations ok" % main())
