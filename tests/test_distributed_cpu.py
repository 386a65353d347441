"""World-size-2 checks of the data-parallel host logic on CPU (gloo): the sharding rules of SURVEY.md 8(e).
  * training: equal shards + ONE all-reduce(avg) of the flat gradient == the single-process gradient of the global
    batch (the oracle plays the UNet here; the CUDA kernels are covered by the -m gpu tests);
  * sampling: `shard_for_rank` partitions the sample set contiguously, in torch.split order, with no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from baddiffusion_b200.model import shard_for_rank
    from oracle import torch_ref as O

    cfg = dict(O.TINY_CONFIG, block_out_channels=(32, 32), layers_per_block=1)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.make_state_dict(cfg, 0).items()}
    B = 4
    g = torch.Generator().manual_seed(0)
    image = torch.randn(B, 3, 32, 32, generator=g).clamp(-1, 1)
    noise = torch.randn(B, 3, 32, 32, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    isp = torch.tensor([True, False, False, True])
    trig, _, acp_all = O.get_trigger("BOX_14", 32), None, None
    targ = O.get_target("CORNER", trig)
    _, alphas, acp = O.beta_tables()
    lo, hi = shard_for_rank(B, rank, world)
    R, x0 = O.poison_blend(image[lo:hi], isp[lo:hi], trig, targ)
    loss = O.p_losses(sd, cfg, alphas, acp, x0, R, t[lo:hi], noise[lo:hi])
    loss.backward()
    flat = torch.cat([v.grad.flatten() for v in sd.values()])   # the flat gradient buffer
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                 # ONE collective per step
    flat /= world
    if rank == 0:
        np.save(os.path.join(out_dir, "dp_grad.npy"), flat.numpy())
        np.save(os.path.join(out_dir, "dp_loss.npy"), np.array([float(loss)]))
    # sampling shards: no collective, contiguous, in order
    n = 11
    mine = list(range(*shard_for_rank(n, rank, world)))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        np.save(os.path.join(out_dir, "shards.npy"), np.array(sum(gathered, [])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_average_matches_global_batch(tmp_path):
    from oracle import torch_ref as O

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cfg = dict(O.TINY_CONFIG, block_out_channels=(32, 32), layers_per_block=1)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.make_state_dict(cfg, 0).items()}
    g = torch.Generator().manual_seed(0)
    image = torch.randn(4, 3, 32, 32, generator=g).clamp(-1, 1)
    noise = torch.randn(4, 3, 32, 32, generator=g)
    t = torch.randint(0, 1000, (4,), generator=g)
    isp = torch.tensor([True, False, False, True])
    trig = O.get_trigger("BOX_14", 32)
    targ = O.get_target("CORNER", trig)
    _, alphas, acp = O.beta_tables()
    R, x0 = O.poison_blend(image, isp, trig, targ)
    loss = O.p_losses(sd, cfg, alphas, acp, x0, R, t, noise)
    loss.backward()
    ref = torch.cat([v.grad.flatten() for v in sd.values()]).numpy()
    got = np.load(tmp_path / "dp_grad.npy")
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    assert np.array_equal(np.load(tmp_path / "shards.npy"), np.arange(11))


def test_shard_for_rank_properties():
    from baddiffusion_b200.model import shard_for_rank

    for n in (0, 1, 7, 16, 2048, 2049):
        for world in (1, 2, 4, 8):
            spans = [shard_for_rank(n, r, world) for r in range(world)]
            flat = [i for lo, hi in spans for i in range(lo, hi)]
            assert flat == list(range(n))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(s for s in sizes if s >= 0) <= max(sizes)  # contiguous, non-increasing tail
            assert all(sizes[i] >= sizes[i + 1] for i in range(world - 1))


def test_bucket_complement_covers_the_gradient_buffer_exactly_once():
    """The overlapped data-parallel step all-reduces the ranges each backward part finished, then `uncovered_ranges`:
    together they must tile [0, n) exactly once (a gap would leave gradients un-averaged, an overlap would average twice)."""
    from baddiffusion_b200.train import uncovered_ranges

    n = 1000
    for covered in ([], [(0, n)], [(100, 300), (300, 450), (700, 900)], [(700, 900), (100, 300)], [(0, 10), (990, n)]):
        pieces = sorted(list(covered) + uncovered_ranges(covered, n))
        assert pieces[0][0] == 0 and pieces[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        assert all(lo < hi for lo, hi in pieces)
