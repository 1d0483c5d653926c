"""`early_stopping` with the reference's interface (/root/reference/libdl/metrics/monitoring.py:4-63), so that the experiment scripts'
`from libdl.metrics import early_stopping, calculate_eval_measures, calculate_mpe_measures_mireval` resolves against this package.
Pure host control flow (one scalar comparison per epoch); nothing here touches the device."""
import math


class early_stopping(object):
    """step(metric) -> True when training should stop: `patience` consecutive epochs without an improvement of more than `min_delta`
    (absolute, or in per cent of the best value with percentage=True) over the best value seen; a NaN metric stops at once;
    patience == 0 disables the check."""

    def __init__(self, mode='min', min_delta=0, patience=10, percentage=False):
        if mode not in {'min', 'max'}:
            raise ValueError('mode ' + mode + ' is unknown!')
        self.mode, self.min_delta, self.patience = mode, min_delta, patience
        self.best, self.num_bad_epochs = None, 0
        margin = (lambda best: best * min_delta / 100) if percentage else (lambda best: min_delta)
        if patience == 0:
            self.is_better = lambda a, best: True
            self.step = lambda a: False
        elif mode == 'min':
            self.is_better = lambda a, best: a < best - margin(best)
        else:
            self.is_better = lambda a, best: a > best + margin(best)

    def step(self, metrics):
        if self.best is None:
            self.best = metrics
            return False
        if math.isnan(metrics):
            return True
        if self.is_better(metrics, self.best):
            self.num_bad_epochs, self.best = 0, metrics
        else:
            self.num_bad_epochs += 1
        return self.num_bad_epochs >= self.patience

    def curr_is_better(self, metrics):
        return self.is_better(metrics, self.best)
