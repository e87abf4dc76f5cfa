"""world_size-2 test of the multi-rank host logic on CPU (gloo): sample-range sharding + one film reduce.
The renderer stand-in is the CPU oracle (the checker), because no GPU is available to the CPU suite; the logic under test is
cudapath.dist (range split, accumulate, reduce), which is exactly what bench.py / a multi-GPU host runs with NCCL."""
import os
import sys
import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
    import torch
    import torch.distributed as dist
    import cudapath
    import orc
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ov = dict(width=24, height=16, spp=5, maxDepth=5)
    env = np.full((8, 16, 3), 1.0, np.float32); env[2, 3] = 40
    s = orc.scene_from_description('straight-hair', scale=0.003, overrides=ov, envmap=env)
    film = torch.zeros((16, 24, 5), dtype=torch.float32)

    def render_range(b, e):
        film.add_(torch.from_numpy(s.render(5, seed=9, sample_begin=b, sample_end=e, threads=1)))

    cudapath.dist.render_sharded(render_range, 5, film, rank, world)
    if rank == 0:
        full = s.render(5, seed=9, threads=1)
        np.savez(out_path, sharded=film.numpy(), full=full)
    dist.barrier()
    dist.destroy_process_group()


def test_sample_range_partition():
    import cudapath
    for total in (0, 1, 5, 64, 1024):
        for world in (1, 2, 3, 8):
            r = [cudapath.dist.sample_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cudapath.dist.sample_range(4, 2, 2)


def test_two_rank_film_reduce_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res.npz')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    assert r['sharded'][..., 4].sum() > 0
    assert np.allclose(r['sharded'], r['full'], rtol=1e-5, atol=1e-6)      # the result is independent of the number of ranks
