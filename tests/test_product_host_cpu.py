"""CPU pins of the PRODUCT's host-side pieces against fixtures generated from the reference (scripts/make_goldens.py):
`baddiffusion_b200.dataset.Backdoor` (dataset.py:526-597,627-655) in both of its modes
(bitmaps on disk / packaged pre-rendered tensors), `train.cosine_lr_lambda` (a23: optimization.py:109-141), the mask
of dataset.py:275-276 and the synthetic data generator the bench uses."""
import os

import numpy as np
import pytest
import torch

from baddiffusion_b200.dataset import Backdoor, SyntheticDataset, get_mask, normalize
from baddiffusion_b200.train import cosine_lr_lambda, uncovered_ranges

T = torch.from_numpy
TRIGGERS = ("BOX_14", "BOX_8", "SM_BOX", "STOP_SIGN_14", "GLASSES", "NONE")
TARGETS = ("HAT", "CAT", "CORNER", "TRIGGER", "SHIFT")


def _check_all(bd, g):
    for S in (32, 256):
        for kind in TRIGGERS:
            got = bd.get_trigger(type=kind, channel=3, image_size=S)
            assert got.dtype == torch.float32 and torch.equal(got, T(g[f"trigger_{kind}_{S}"])), (kind, S)
        base = bd.get_trigger(type="BOX_14", channel=3, image_size=S)
        for kind in TARGETS:
            got = bd.get_target(type=kind, trigger=base)
            assert torch.equal(got, T(g[f"target_{kind}_{S}"])), (kind, S)


def test_backdoor_packaged_assets_bit_exact(golden, tmp_path):
    """No static/ directory: the bitmap-derived tensors come from assets/backdoor_assets.npz."""
    _check_all(Backdoor(root="datasets", static_root=str(tmp_path)), golden("backdoor_tensors"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/static"), reason="reference bitmaps not mounted")
def test_backdoor_bitmap_mode_bit_exact(golden):
    """static/*.png on disk: PIL decode + Resize + normalize + pad, the reference's own path (dataset.py:420-497)."""
    _check_all(Backdoor(root="datasets", static_root="/root/reference"), golden("backdoor_tensors"))


def test_backdoor_other_sizes_and_errors():
    bd = Backdoor(root="datasets")
    for kind, k, val in (("BOX_18", 18, 0.0), ("BOX_11", 11, 0.0), ("BOX_4", 4, 0.0), ("BIG_BOX", 18, 1.0),
                         ("XSM_BOX", 11, 1.0), ("XXSM_BOX", 8, 1.0), ("XXXSM_BOX", 4, 1.0)):
        t = bd.get_trigger(type=kind, channel=3, image_size=32)
        assert int((t > -1).sum()) == 3 * k * k and float(t.max()) == val
        assert torch.equal(t[:, 32 - 2 - k: 32 - 2, 32 - 2 - k: 32 - 2], torch.full((3, k, k), val))
    t1 = bd.get_trigger(type="BOX_14", channel=1, image_size=28)   # MNIST geometry
    assert t1.shape == (1, 28, 28) and int((t1 > -1).sum()) == 196
    with pytest.raises(ValueError):
        bd.get_trigger(type="NOPE", channel=3, image_size=32)
    with pytest.raises(NotImplementedError):
        bd.get_target(type="NOPE", trigger=t1)


def test_mask_and_normalize(golden):
    g = golden("backdoor_tensors")
    trig = T(g["trigger_BOX_14_32"])
    m = get_mask(trig)
    assert int((m == 0).sum()) == 588 and set(m.unique().tolist()) == {0, 1}
    x = torch.tensor([0.0, 0.5, 1.0])
    y = normalize(x, vmin_in=0.0, vmax_in=1.0, vmin_out=-1.0, vmax_out=1.0)     # quirk Q7: [0,1] -> [-1, 1-2e-5]
    assert float(y[0]) == -1.0 and abs(float(y[2]) - (1.0 - 2e-5)) < 1e-6


def test_cosine_lr_product(golden):
    lrs = golden("cosine_lr")["lrs"]
    mine = np.array([2e-4 * cosine_lr_lambda(i, 500, 2000) for i in range(2000)])
    assert np.abs(mine - lrs).max() < 1e-12
    assert cosine_lr_lambda(0, 500, 2000) == 0.0 and cosine_lr_lambda(500, 500, 2000) == 1.0
    assert cosine_lr_lambda(2000, 500, 2000) < 1e-12 and cosine_lr_lambda(5000, 500, 2000) >= 0.0


def test_synthetic_dataset_protocol():
    """image = randn(gen(seed + index)).clamp(-1, 1); every round(1/rate)-th sample is poisoned."""
    ds = SyntheticDataset(32, 3, poison_rate=0.1, seed=5)
    b = ds.batch(20, index=2, pin=False)
    ref = torch.randn(20, 3, 32, 32, generator=torch.Generator().manual_seed(7)).clamp(-1, 1)
    assert torch.equal(b.image, ref) and b.is_poison.tolist() == [1 if i % 10 == 0 else 0 for i in range(20)]
    assert torch.equal(ds.trigger, Backdoor(root="x").get_trigger(type="BOX_14", channel=3, image_size=32))


def test_uncovered_ranges():
    assert uncovered_ranges([(10, 20), (0, 5)], 30) == [(5, 10), (20, 30)]
    assert uncovered_ranges([], 7) == [(0, 7)] and uncovered_ranges([(0, 7)], 7) == []


def test_pndm_scheduler_bookkeeping_matches_reference(golden):
    """PNDMScheduler.set_timesteps (prk / plms timestep lists) == the reference's, and the host-side step
    bookkeeping (modes, history slots, formula-(9) scalars) reproduces the reference's step outputs when the row is
    evaluated with the kernel's arithmetic restated in torch fp32 (the CUDA kernel itself is checked in the GPU suite)."""
    import numpy as np
    import torch

    from baddiffusion_b200 import schedulers as SCH

    g = golden("pndm")
    k16, k13, k124 = (torch.tensor(v, dtype=torch.float32) for v in (1 / 6, 1 / 3, 1 / 24))
    for skip in (0, 1):
        for nsteps in (50, 20):
            tag = f"steps_skip{skip}_{nsteps}"
            s = SCH.PNDMScheduler(skip_prk_steps=bool(skip))
            s.set_timesteps(nsteps)
            assert np.array_equal(s.timesteps.numpy(), g[f"{tag}/timesteps"])
            x = torch.from_numpy(g[f"{tag}/x0"])
            acc = torch.zeros_like(x)
            cur = torch.zeros_like(x)
            ets = [torch.zeros_like(x) for _ in range(4)]
            for i in range(g[f"{tag}/eps"].shape[0]):
                eps = torch.from_numpy(g[f"{tag}/eps"][i])
                row = s.next_row(int(s.timesteps[i]))
                mode, push = int(row[0]), int(row[5])
                sl = [int(v) for v in row[6:10]]
                e = lambda k: eps if sl[k] == push else ets[sl[k]]
                S, M = x, eps
                if mode == SCH.PNDM_PRK0:
                    acc, cur = k16 * eps, x
                elif mode == SCH.PNDM_PRK12:
                    acc, S = acc + k13 * eps, cur
                elif mode == SCH.PNDM_PRK3:
                    M, S = acc + k16 * eps, cur
                elif mode == SCH.PNDM_PLMS_FIRST:
                    cur = x
                elif mode == SCH.PNDM_PLMS_SECOND:
                    M, S = (eps + ets[sl[0]]) / 2, cur
                elif mode == SCH.PNDM_PLMS2:
                    M = (3 * e(0) - ets[sl[1]]) / 2
                elif mode == SCH.PNDM_PLMS3:
                    M = (23 * e(0) - 16 * ets[sl[1]] + 5 * ets[sl[2]]) / 12
                else:
                    M = k124 * (55 * e(0) - 59 * ets[sl[1]] + 37 * ets[sl[2]] - 9 * ets[sl[3]])
                if push >= 0:
                    ets[push] = eps
                x = row[1] * S - row[2] * M / row[3]
                assert torch.equal(x, torch.from_numpy(g[f"{tag}/out"][i])), (tag, i, mode)
