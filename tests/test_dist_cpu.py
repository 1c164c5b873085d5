"""world_size-2 gloo test of the multi-rank host logic (sharding, counter reduction, stop-rule replay)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparc_ldpc_b200 import dist as sd
    total = 11
    mine = sd.shard_indices(total)
    # every rank "decodes" its shard: block g has g % 3 bit errors
    errs = np.array([[g % 3, g] for g in mine], dtype=np.float64)
    counts = sd.allreduce_counts(np.array([errs[:, 0].sum(), len(mine)], dtype=np.int64))
    rows = sd.gather_rows(errs, total)
    used = sd.replay_stop_rule(rows[:, 0] != 0, min_errors=4, max_blocks=100)
    q.put((rank, mine.tolist(), counts.tolist(), rows[:, 1].tolist(), used))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert out[0][1] == [0, 2, 4, 6, 8, 10] and out[1][1] == [1, 3, 5, 7, 9]
    total_err = sum(g % 3 for g in range(11))
    for r in out:
        assert r[2] == [total_err, 11]
        assert r[3] == list(map(float, range(11)))
        # blocks 1,2,4,5 are the first four error blocks -> the reference loop stops after block index 5
        assert r[4] == 6


def test_stop_rule_replay_edges():
    from sparc_ldpc_b200 import dist as sd
    assert sd.replay_stop_rule([0, 0, 0], 1, 2) == 2          # MAX_BLOCKS cut
    assert sd.replay_stop_rule([1, 0, 0], 1, 10) == 1
    assert sd.replay_stop_rule([0, 0], 5, 10) == 2            # ran out of drawn blocks
    assert sd.shard_indices(5, 1, 2).tolist() == [1, 3]


def _mc_worker(rank, world, port, q):
    """The parity-mode Monte-Carlo loop (_mc of sparc_ldpc_b200/sparc_ldpc.py) with a host-only 'decoder': all
    ranks walk the same RNG stream, decode blocks j mod world of every chunk, exchange the rows and replay the stop
    rule -- rows, block count and final RNG state must equal the one-rank run."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparc_ldpc_b200 import dist as sd, sparc_ldpc as S
    calls = []

    def run(group):
        rng = np.random.RandomState(123)

        def draw_block():
            return rng.randn(5), rng.randint(0, 2, 3)

        def decode_blocks(blocks):       # row = (array[2], array[1], scalar): the shape of a _pair_driver row
            calls.append(len(blocks))
            return [(np.array([b[0].sum(), b[0].max()]), np.array([float(b[1].sum())]), float(b[0][0] > 0.3)) for b in blocks]

        rows = S._mc(rng, draw_block, decode_blocks, lambda row: row[2] != 0, 4, 40, 7, group)
        return rows, rng.get_state()

    single, st1 = run(None)
    n_single = sum(calls)
    del calls[:]
    shard, st2 = run(True)
    n_shard = sum(calls)
    same = len(single) == len(shard) and all(
        np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] for a, b in zip(single, shard))
    same_rng = st1[0] == st2[0] and np.array_equal(st1[1], st2[1]) and st1[2:] == st2[2:]
    # a flow with scalar rows (sim_ldpc) and a chunk smaller than the world (one rank idle)
    flat = sd.decode_sharded(lambda bl: [float(x) * 2 for x in bl], [3.0], True)
    q.put((rank, same, same_rng, len(single), n_single, n_shard, flat))
    dist.destroy_process_group()


def test_two_rank_monte_carlo_loop_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_mc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for r in out:
        assert r[1] and r[2], r                 # identical rows and identical final RNG state
        assert r[6] == [6.0]
    # the two ranks together decoded exactly the blocks one rank decodes alone
    assert out[0][5] + out[1][5] == out[0][4] and out[0][5] > 0 and out[1][5] > 0
