"""Learning-rate schedules of `pt/solver/build.py:build_lr_scheduler` as pure functions of the iteration.

The reference steps a torch `_LRScheduler` object once per iteration (a d2 hook); on this path the optimizer is one
fused kernel over the flat arenas that takes the learning rate as an argument, so the schedule is evaluated in
closed form from `PTrainer.iter` (which also makes resuming trivial: no scheduler state to restore).
  WarmupMultiStepLR          detectron2 v0.5 solver/lr_scheduler.py (configs/pt/final_c2f.yaml:6)
  WarmupCosineLR             detectron2 v0.5 solver/lr_scheduler.py
  WarmupTwoStageMultiStepLR  pt/solver/lr_scheduler.py:21-66 (factor_list indexed by the number of milestones passed)
"""
import math
from bisect import bisect_right


def get_warmup_factor_at_iter(method, it, warmup_iters, warmup_factor):
    """detectron2 v0.5 `_get_warmup_factor_at_iter`."""
    if it >= warmup_iters:
        return 1.0
    if method == "constant":
        return warmup_factor
    if method == "linear":
        alpha = it / warmup_iters
        return warmup_factor * (1 - alpha) + alpha
    raise ValueError("Unknown warmup method: {}".format(method))


def lr_at_iter(cfg, it):
    """Learning rate of iteration `it` for cfg.SOLVER.LR_SCHEDULER_NAME (same dispatch / errors as
    `pt/solver/build.py:27-60`)."""
    s = cfg.SOLVER
    name = s.LR_SCHEDULER_NAME
    warm = get_warmup_factor_at_iter(s.WARMUP_METHOD, it, s.WARMUP_ITERS, s.WARMUP_FACTOR)
    if name == "WarmupMultiStepLR":
        steps = list(s.STEPS)
        if steps != sorted(steps):
            raise ValueError("Milestones should be a list of increasing integers. Got {}".format(steps))
        return s.BASE_LR * warm * s.GAMMA ** bisect_right(steps, it)
    if name == "WarmupCosineLR":
        return s.BASE_LR * warm * 0.5 * (1.0 + math.cos(math.pi * it / s.MAX_ITER))
    if name == "WarmupTwoStageMultiStepLR":
        steps, factors = list(s.STEPS), list(s.FACTOR_LIST)
        if steps != sorted(steps):
            raise ValueError("Milestones should be a list of increasing integers. Got {}".format(steps))
        if len(steps) + 1 != len(factors):
            raise ValueError("Length of milestones should match length of factor_list.")
        return s.BASE_LR * warm * factors[bisect_right(steps, it)]
    raise ValueError("Unknown LR scheduler: {}".format(name))
