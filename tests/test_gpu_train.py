"""The training loop around the hot path (systems/base.py:100-116 on_train_batch_start -> update_module_step,
training_step, optimizer + SequentialLR stepped per batch): `train.NeusTrainer` / `train.SplitTrainer` FROM SCRATCH --
no hand-set occupancy grid, level mask, stage or finite-difference eps: everything comes from `update_step`."""
import math

import pytest
import torch

from rise_sdf_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def test_neus_training_from_scratch():
    from rise_sdf_b200.neus import NeuSModel, neus_blender_config
    from rise_sdf_b200.train import NeusTrainer
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    m = NeuSModel(neus_blender_config()).cuda().train()
    tr = NeusTrainer(m)
    assert float(m.occupancy_grid.binaries.float().sum()) == 0          # nothing has been set up by hand
    batch = [t.cuda() for t in syn.training_rays(1024, seed=1)]
    losses, lrs = [], []
    for it in range(34):                                                # occupancy updates at steps 0, 16, 32
        lrs.append(tr.opt.param_groups[0]["lr"])
        loss, out = tr.step(*batch)
        assert torch.isfinite(loss), it
        losses.append(float(loss))
        if it == 0:
            assert 0.05 < float(m.occupancy_grid.binaries.float().mean()) < 0.99      # warm-up branch filled the grid
            assert int(out["num_samples"]) > 10000
    assert tr.global_step == 34 and m.cos_anneal_ratio == pytest.approx(33 / 20000)
    # configs/neus-blender.yaml:104-119: LinearLR 0.01 -> 1 over 500 steps on lr 0.01 (variance group 0.001)
    assert lrs[0] == pytest.approx(0.01 * 0.01) and lrs[33] == pytest.approx(0.01 * (0.01 + 0.99 * 33 / 500))
    assert tr.opt.param_groups[2]["lr"] == pytest.approx(0.001 * (0.01 + 0.99 * 34 / 500))
    assert sum(losses[-5:]) / 5 < sum(losses[:5]) / 5                   # the same batch 34 times: it must fit it
    assert float(tr.opt.state[m.variance.variance]["step"]) == 34.0


def test_split_training_from_scratch_stage_switch_and_emitter_group():
    from rise_sdf_b200.split_mixed_occ import SplitMixedOCCModel, split_mixed_occ_config
    from rise_sdf_b200.train import SplitTrainer
    torch.manual_seed(0)
    torch.cuda.manual_seed(0)
    cfg = split_mixed_occ_config()
    cfg["light"]["envlight_config"]["base_res"] = 64
    m = SplitMixedOCCModel(cfg).cuda().train()
    tr = SplitTrainer(m)
    batch = [t.cuda() for t in syn.training_rays(512, seed=2)]
    base0 = m.emitter.base.detach().clone()
    for it in range(3):
        loss, out = tr.step(*batch)
        assert torch.isfinite(loss) and m.stage == 0
    # stage 0: the progressive mask starts at level 6 (start_level), eps = 2 r / (32 * pls^5), emitter off the graph
    assert int(m.geometry.encoding.encoding.mask.sum()) == 12
    assert m.geometry._finite_difference_eps == pytest.approx(3.0 / (32 * 1.447269237440378 ** 5))
    assert torch.equal(m.emitter.base, base0)
    assert float(tr.opt.state[m.emitter.base]["step"]) == 0.0 and 0 < float(tr.opt.state[m.variance.variance]["step"]) == 3.0
    tr.global_step = 10000                                              # split_sum_kick_in_step
    lr = tr.opt.param_groups[3]["lr"]
    loss, out = tr.step(*batch)
    assert m.stage == 1 and "comp_rgb_phys_full" in out and torch.isfinite(loss)
    assert int(m.geometry.encoding.encoding.mask.sum()) == 2 * min(6 + (10000 - 6000) // 500, 16)
    assert not torch.equal(m.emitter.base, base0) and float(tr.opt.state[m.emitter.base]["step"]) == 1.0
    # first emitter update after 10 000 idle steps: |delta| ~ lr (bias-corrected), not ~3.2 lr
    step = (m.emitter.base - base0).abs()
    moved = step[step > 0]
    assert moved.numel() > 100 and float(moved.max()) <= 1.01 * lr and float(moved.median()) > 0.5 * lr
