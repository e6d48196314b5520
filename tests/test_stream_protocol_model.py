"""Executable model of the stream program's hand-off (csrc/program_stream.cuh): tagged words in rotating rows.

Op i of a run publishes its outputs as (fp16 | tag) words into row i % 4, tag = (base + i) % 65535 + 1, base advancing
by n_ops per run (mod 65535); a consumer polls the row of its source op j in {i-1, i-2, i-3} until every word carries
tag(base, j).  Nothing is ever cleared, so safety rests on one property: at the moment a consumer may START polling
(any time after the row's previous use), no word of the columns it reads may ALREADY carry the awaited tag unless op j
of THIS run wrote it.  The model replays many runs of random programs (different widths per op, so narrow ops leave
stale words of older, wider ops behind them), across the wrap-around of the tag base, and checks exactly that - plus
the row-reuse argument (the writer of row r in op i + 4 cannot run before every reader of op i's row is done, because
finishing op i + 3 needs everybody's op i + 2 outputs)."""
import random

ROWS = 4


def tag(base, op):
    return (base + op) % 65535 + 1


def run_model(n_ops, widths, sources, runs, base0):
    rows = [[0] * max(widths) for _ in range(ROWS)]      # zero-initialised once, never cleared
    base = base0
    for _ in range(runs):
        for i in range(n_ops):
            j = sources[i]
            if j is not None:
                # the consumer's poll of op j's row must succeed now (op j of this run wrote it) ...
                want = tag(base, j)
                assert all(rows[j % ROWS][c] == want for c in range(widths[j])), (i, j)
            # ... and BEFORE op i publishes, nobody polling for op i's tag may see it anywhere in the row
            mine = tag(base, i)
            assert mine not in rows[i % ROWS], f"stale word already carries the tag of op {i} (base {base})"
            for c in range(widths[i]):
                rows[i % ROWS][c] = mine
        base = (base + n_ops) % 65535
    return True


def test_tags_never_alias_across_runs_and_wraparound():
    rng = random.Random(7)
    for trial in range(40):
        n_ops = rng.choice([1, 2, 3, 4, 5, 7, 8, 16, 128, 255])
        widths = [rng.choice([16, 64, 256, 1024]) for _ in range(n_ops)]
        sources = [None if i == 0 or rng.random() < 0.1 else rng.randint(max(0, i - 3), i - 1) for i in range(n_ops)]
        # start close to the wrap of the tag base so that the modulo is exercised
        base0 = (65535 - rng.randint(0, 3 * n_ops)) % 65535
        assert run_model(n_ops, widths, sources, runs=rng.randint(5, 40), base0=base0)


def test_a_full_tag_period_of_a_decode_step():
    """128 ops per run (the Llama-3-8B step): 65535 / gcd(128, 65535) = 65535 runs until the base repeats; a stale word is
    at most one run (128 ops) old when its column is rewritten, so aliasing would need two ops 65535 apart in the global
    op count writing the same row - check the arithmetic that rules it out instead of simulating 8 M ops."""
    n_ops = 128
    seen = {}
    base = 0
    for run in range(3):
        for i in range(n_ops):
            t = tag(base, i)
            key = (i % ROWS, t)
            # within any window of 2 runs no (row, tag) pair repeats
            assert key not in seen or (run * n_ops + i) - seen[key] >= 65535, key
            seen[key] = run * n_ops + i
        base = (base + n_ops) % 65535
    # n_ops must stay below the tag period (checked at creation: n < 60000)
    assert n_ops < 60000
