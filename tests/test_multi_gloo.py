"""World-size-2 gloo test (CPU) of the multi-rank logic bench.py uses for --gpus N: distinct sequences per rank,
barrier, max-over-ranks timing, whole-job aggregation."""
import os

import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from rebvo_b200 import multi
    dist = multi.init("gloo")
    assert multi.rank_info() == (rank, rank, world)
    seed = multi.stream_seed(rank)
    dist.barrier()
    t_ms = [100.0 + 50.0 * rank, 10.0 * (world - rank)]
    mx = multi.max_over_ranks(dist, t_ms)
    fps = multi.aggregate_fps(64 * 8, world, mx[0])
    q.put((rank, seed, mx, fps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_max_and_aggregate():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] != out[1][1], "ranks must replay different sequences"
    for _, _, mx, fps in out:
        assert mx == [150.0, 20.0]
        assert abs(fps - 2 * 512 / 0.150) < 1e-6
