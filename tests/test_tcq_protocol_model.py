"""Executable model of the small-M tensor-core kernel's pipeline protocol (csrc/gemm_tc.cu: gemm_tcq_kernel), run on
the CPU.  The kernel's six kinds of warps talk through mbarrier rings only - packed stages (Q-TMA -> both producer
teams), A stages (team t -> MMA warp t), activation stages (X-TMA -> MMA warps), two TMEM accumulator buffers (MMA warps
-> epilogue) - and every ring index / phase-parity expression of the device code is restated here verbatim.  Random
interleavings of the roles (and of the asynchronous agents: TMA landings, tensor-core execution and its commits) check

  * liveness: every role finishes (no lost wake-up, no wait on a parity that never comes);
  * safety: a stage is never overwritten before its last reader has read it, and every reader sees exactly the step it
    expects (packed stage d -> both teams, A stage of k-step s -> the MMA of k-step s, X stage likewise), also across ring
    wrap-arounds, several segments per CTA (TMEM buffer recycling) and one-pair ranges;
  * the accumulators the epilogue drains hold exactly the k-steps of the segment, once each.

The model knows nothing about CUDA: it pins the index arithmetic, which is where such kernels break.  (The real
kernel's own tests are tests/test_gpu_parity.py::test_small_m_*.)"""
import random

import pytest


class Mbar:
    """mbarrier with a pending-arrival count per phase; wait(parity) passes once the latest phase of that parity is
    complete (a fresh barrier lets parity 1 through: the `phase ^ 1` idiom of the producers)."""

    def __init__(self, count):
        self.count, self.arrivals, self.done = count, 0, 0   # done = completed phases

    def arrive(self):
        self.arrivals += 1
        assert self.arrivals <= self.count, "more arrivals than the barrier was initialised for"
        if self.arrivals == self.count:
            self.arrivals, self.done = 0, self.done + 1

    def passed(self, parity):
        return (self.done & 1) != parity


class Sim:
    def __init__(self, NS, NX, NQ, ranges, KP, seed):
        self.NS, self.NX, self.NQ, self.KP = NS, NX, NQ, KP
        self.t_begin, self.t_end = ranges
        self.rng = random.Random(seed)
        self.full = [Mbar(8) for _ in range(NS)]          # 8 producer warps of one team (modelled as 1 agent x 8 arrivals)
        self.empty = [Mbar(1) for _ in range(NS)]
        self.xfull = [Mbar(2) for _ in range(NX)]         # expect_tx arrival + "bytes landed"
        self.xempty = [Mbar(1) for _ in range(NX)]
        self.qfull = [Mbar(2) for _ in range(NQ)]
        self.qempty = [Mbar(16) for _ in range(NQ)]       # both teams
        self.tmem_full = [Mbar(2) for _ in range(2)]      # one commit per MMA warp
        self.tmem_empty = [Mbar(128) for _ in range(2)]
        self.q_data = [None] * NQ                          # content tags: what a stage currently holds
        self.a_data = [None] * NS
        self.x_data = [None] * NX
        self.acc = [[[] for _ in range(2)] for _ in range(2)]   # [buf][team] -> accumulated k-steps
        self.async_ops = []                                # pending asynchronous actions (TMA landings, MMA execution)
        self.drained = []                                  # (segment, steps) seen by the epilogue

    # ---- segment walk shared by all roles (device: nt / d0 / d1 from t)
    def segments(self):
        t, out = self.t_begin, []
        while t < self.t_end:
            nt = t // self.KP
            d0 = t - nt * self.KP
            d1 = self.KP if self.KP - d0 < self.t_end - t else d0 + (self.t_end - t)
            out.append((nt, d0, d1))
            t += d1 - d0
        return out

    # ---- roles as generators: `yield cond` blocks until cond() is true
    def q_tma(self):
        qs, qph = 0, 0
        for nt, d0, d1 in self.segments():
            for d in range(d0, d1):
                yield lambda qs=qs, qph=qph: self.qempty[qs].passed(qph ^ 1)
                self.qfull[qs].arrive()                                    # arrive.expect_tx
                self.async_ops.append(("land_q", qs, (nt, d)))
                qs += 1
                if qs == self.NQ:
                    qs, qph = 0, qph ^ 1

    def x_tma(self):
        xs, xph = 0, 0
        for nt, d0, d1 in self.segments():
            for s in range(2 * d0, 2 * d1):
                yield lambda xs=xs, xph=xph: self.xempty[xs].passed(xph ^ 1)
                self.xfull[xs].arrive()
                self.async_ops.append(("land_x", xs, (nt, s)))
                xs += 1
                if xs == self.NX:
                    xs, xph = 0, xph ^ 1

    def producer(self, team):
        steps = [(nt, d) for nt, d0, d1 in self.segments() for d in range(d0, d1)]
        npairs = len(steps)
        qs, stage, qph, phase = 0, team, 0, 0
        cur = None
        if npairs > 0:
            yield lambda: self.qfull[0].passed(0)
            cur = self.q_data[0]
        for i in range(npairs):
            qs_n = 0 if qs + 1 == self.NQ else qs + 1
            qph_n = qph ^ 1 if qs + 1 == self.NQ else qph
            nxt = None
            if i + 1 < npairs:
                yield lambda qs_n=qs_n, qph_n=qph_n: self.qfull[qs_n].passed(qph_n)
                nxt = self.q_data[qs_n]                                    # fetch: LDS of the next stage
            yield lambda stage=stage, phase=phase: self.empty[stage].passed(phase ^ 1)
            assert cur == steps[i], f"team {team}: packed stage holds {cur}, expected {steps[i]}"
            nt, d = steps[i]
            self.a_data[stage] = (nt, 2 * d + team)                       # dequantised tile of k-step 2 d + team
            for _ in range(8):
                self.full[stage].arrive()
                self.qempty[qs].arrive()
            stage += 2
            if stage >= self.NS:
                stage, phase = stage - self.NS, phase ^ 1
            cur, qs, qph = nxt, qs_n, qph_n

    def mma(self, mw):
        stage, xs, phase, xph = mw, mw, 0, 0
        for it, (nt, d0, d1) in enumerate(self.segments()):
            buf, use = it & 1, (it >> 1) & 1
            yield lambda buf=buf, use=use: self.tmem_empty[buf].passed(use ^ 1)
            for d in range(d0, d1):
                yield lambda xs=xs, xph=xph: self.xfull[xs].passed(xph)
                yield lambda stage=stage, phase=phase: self.full[stage].passed(phase)
                # issue: the tensor core reads the operands LATER (asynchronously), then the commits arrive
                self.async_ops.append(("mma", mw, buf, stage, xs, (nt, 2 * d + mw), d == d0))
                stage += 2
                if stage >= self.NS:
                    stage, phase = stage - self.NS, phase ^ 1
                xs += 2
                if xs >= self.NX:
                    xs, xph = xs - self.NX, xph ^ 1
            self.async_ops.append(("commit_acc", mw, buf))

    def epilogue(self):
        for it, (nt, d0, d1) in enumerate(self.segments()):
            buf, use = it & 1, (it >> 1) & 1
            yield lambda buf=buf, use=use: self.tmem_full[buf].passed(use)
            got = sorted(self.acc[buf][0] + self.acc[buf][1])
            want = [(nt, s) for s in range(2 * d0, 2 * d1)]
            assert got == want, f"segment {it}: accumulators hold {got}, expected {want}"
            self.drained.append((it, got))
            for _ in range(128):
                self.tmem_empty[buf].arrive()

    # ---- asynchronous agents: per-queue FIFO order (TMA per ring, tensor core per issuing warp), random across queues
    def step_async(self):
        if not self.async_ops:
            return False
        # the tensor core executes one warp's MMAs / commits in issue order; TMA landings of one ring may complete in
        # any order relative to other rings - pick any op that is first of its (kind, owner) queue
        firsts, seen = [], set()
        for i, op in enumerate(self.async_ops):
            key = ("tc", op[1]) if op[0] in ("mma", "commit_acc") else (op[0], op[1])
            if key not in seen:
                seen.add(key)
                firsts.append(i)
        op = self.async_ops.pop(self.rng.choice(firsts))
        if op[0] == "land_q":
            _, qs, tag = op
            self.q_data[qs] = tag
            self.qfull[qs].arrive()
        elif op[0] == "land_x":
            _, xs, tag = op
            self.x_data[xs] = tag
            self.xfull[xs].arrive()
        elif op[0] == "mma":
            _, mw, buf, stage, xs, want, first = op
            assert self.a_data[stage] == want, f"MMA {mw}: A stage {stage} holds {self.a_data[stage]}, expected {want}"
            assert self.x_data[xs] == want, f"MMA {mw}: X stage {xs} holds {self.x_data[xs]}, expected {want}"
            if first:
                self.acc[buf][mw] = []
            self.acc[buf][mw].append(want)
            self.empty[stage].arrive()        # tcgen05.commit -> empty[stage], xempty[xs]
            self.xempty[xs].arrive()
        else:
            _, mw, buf = op
            self.tmem_full[buf].arrive()
        return True

    def run(self):
        roles = {"q": self.q_tma(), "x": self.x_tma(), "p0": self.producer(0), "p1": self.producer(1),
                 "m0": self.mma(0), "m1": self.mma(1), "e": self.epilogue()}
        waiting = {}
        for name, gen in list(roles.items()):
            try:
                waiting[name] = next(gen)
            except StopIteration:
                del roles[name]
        idle = 0
        while roles:
            progressed = False
            choices = list(roles) + ["async"] * 3
            self.rng.shuffle(choices)
            for name in choices:
                if name == "async":
                    progressed |= self.step_async()
                    continue
                if name in roles and waiting[name]():
                    try:
                        waiting[name] = roles[name].send(None)
                    except StopIteration:
                        del roles[name]
                    progressed = True
            idle = 0 if progressed else idle + 1
            assert idle < 3, f"deadlock: {sorted(roles)} blocked, {len(self.async_ops)} async ops pending"
        while self.step_async():
            pass
        assert len(self.drained) == len(self.segments())


CONFIGS = [
    # NS, NX, NQ   (BT <= 16: 4 / 16 / 14; BT = 64: 4 / 8 / 11; BT = 128: 4 / 4 / 11) + small rings that wrap constantly
    (4, 16, 14), (4, 8, 11), (4, 4, 11), (4, 2, 2), (2, 2, 3),
]
RANGES = [
    # (t_begin, t_end), KP: one segment, straddling segments, whole tiles, three segments (buffer recycling), one pair
    ((0, 32), 32), ((20, 75), 32), ((64, 128), 32), ((30, 100), 32), ((7, 8), 9), ((5, 40), 9), ((0, 3), 1),
]


@pytest.mark.parametrize("NS,NX,NQ", CONFIGS)
@pytest.mark.parametrize("rng_range,KP", RANGES)
def test_pipeline_protocol_random_interleavings(NS, NX, NQ, rng_range, KP):
    for seed in range(6):
        Sim(NS, NX, NQ, rng_range, KP, seed).run()


def test_model_detects_a_missing_release():
    """The model is not vacuous: without the producers' wait for `empty` an A stage is overwritten before its MMA ran."""

    class Broken(Sim):
        def producer(self, team):
            for cond in Sim.producer(self, team):
                yield cond if "empty" not in cond.__code__.co_names else (lambda: True)

    caught = 0
    for seed in range(20):
        try:
            Broken(2, 16, 14, (0, 64), 32, seed).run()
        except AssertionError:
            caught += 1
    assert caught > 0
