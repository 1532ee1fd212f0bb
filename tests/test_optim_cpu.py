"""Host-side logic of the optimiser row, CPU only: the warm-up + multi-step schedule against torch's own
chained scheduler (which is what the reference instantiates, engine/my_schedulers.py:50-70), the shard partition of the
sharded optimiser, and the loss scaling quirk."""
import torch

from presight_b200 import optim, schedulers


def test_warmup_multistep_matches_torch_chained_scheduler():
    cfg = schedulers.WarmupMultiStepSchedulerConfig(max_steps=400, milestones=(100, 200, 300), warmup_steps=40)
    sch = cfg.setup()
    p = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.Adam([p], lr=1e-2, eps=1e-15, weight_decay=1e-5)
    s = sch.get_scheduler(opt, 1e-2)
    for step in range(400):
        assert abs(opt.param_groups[0]["lr"] - 1e-2 * sch.lr_factor(step)) < 1e-12, step
        opt.step()
        s.step()
    # the shipped schedule (method_configs.py:116-119): 1 % at step 0, full rate after a tenth of the run, x0.33 per quarter
    ps = schedulers.presight_scheduler(100000).setup()
    assert ps.lr_factor(0) == 0.01 and ps.lr_factor(10000) == 1.0 and abs(ps.lr_factor(5000) - 0.505) < 1e-12
    assert abs(ps.lr_factor(25000) - 0.33) < 1e-12 and abs(ps.lr_factor(99999) - 0.33 ** 3) < 1e-12
    # config.gamma is ignored, as in the reference (the factor is hard-coded)
    assert schedulers.WarmupMultiStepSchedulerConfig(gamma=0.5, milestones=(10,), warmup_steps=0).setup().lr_factor(10) == 0.33


def test_shard_bounds_partition():
    for numel in (1, 7, 64, 1000, (1 << 22) * 2 * 16, 12345677):
        for world in (1, 2, 3, 8):
            covered = 0
            pers = set()
            for r in range(world):
                lo, hi, per = optim.shard_bounds(numel, r, world)
                assert lo == min(r * per, numel) and lo <= hi <= numel and hi - lo <= per and per % 4 == 0
                covered += hi - lo
                pers.add(per)
            assert covered == numel and len(pers) == 1 and per * world >= numel


def test_scale_loss_is_the_grad_scaler_quirk():
    x = torch.tensor(3.0, requires_grad=True)
    optim.scale_loss(x * x).backward()
    assert float(x.grad) == 6.0 * 1024.0
