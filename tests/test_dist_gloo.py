"""N>1 host logic on CPU with the gloo backend, world_size 2: sharding, setup broadcast, length gather, max-over-ranks."""
import os
import socket

import torch
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
    from chatttsplus_b200 import dist as D
    lo, hi = D.shard_range(7, world, rank)
    spk = torch.arange(768, dtype=torch.float32) if rank == 0 else torch.zeros(768)
    D.broadcast_setup(spk, 0)
    local_lens = [10 * rank + i for i in range(hi - lo)]
    all_lens = D.gather_lengths(local_lens)
    t = D.max_over_ranks(1.0 + rank)
    q.put((rank, lo, hi, float(spk[5]), all_lens, t, D.rank_seed(1234, lo)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_broadcast_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, l0, t0, sd0), (r1, lo1, hi1, s1, l1, t1, sd1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 3, 3, 7)                 # contiguous cover of 7 utterances
    assert s0 == s1 == 5.0                                      # speaker embedding reached rank 1
    assert l0 == l1 == [0, 1, 2, 10, 11, 12, 13]                # every rank knows every length, shard sizes differ
    assert t0 == t1 == 2.0                                      # max over ranks
    assert sd0 != sd1


def test_shard_range_covers_everything():
    from chatttsplus_b200.dist import shard_range
    for n in (1, 7, 32, 256):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_utterance_uniforms_do_not_depend_on_the_sharding():
    """Per-utterance RNG streams: the uniforms utterance g consumes are the same whether the job runs on 1, 2, 4 or 8
    ranks and whatever the slice size."""
    from chatttsplus_b200.dist import shard_range, utterance_uniforms
    n, max_new, nq = 16, 5, 4
    whole = utterance_uniforms(99, 0, n, max_new, nq)
    assert whole.shape == (max_new, n * nq)
    for world in (2, 4, 8):
        for slice_size in (1, 2, 3):
            parts = []
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                for a in range(lo, hi, slice_size):
                    parts.append(utterance_uniforms(99, a, min(hi, a + slice_size), max_new, nq))
            assert torch.equal(torch.cat(parts, 1), whole)
    assert not torch.equal(utterance_uniforms(100, 0, 2, max_new, nq), whole[:, : 2 * nq])
