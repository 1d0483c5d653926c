"""Early stopping for the training scripts, API-compatible with the reference class (/root/reference/libdl/metrics/monitoring.py:4-63), so
that `from libdl.metrics import early_stopping, calculate_eval_measures, calculate_mpe_measures_mireval` resolves against this package.
Host control flow only: one scalar comparison per epoch, nothing touches the device.

Behaviour (pinned by tests/golden/ext_golden.npz against the reference class's decisions):
  * the first value becomes the best so far and never stops training;
  * a NaN value stops at once;
  * a value counts as an improvement when it beats the best by more than `min_delta` — an absolute margin, or (percentage=True)
    that many per cent of the best value;
  * `patience` epochs without improvement in a row stop training; patience == 0 switches the check off."""
import math


class early_stopping(object):
    def __init__(self, mode='min', min_delta=0, patience=10, percentage=False):
        if mode not in ('min', 'max'):
            raise ValueError('mode ' + mode + ' is unknown!')
        self.mode = mode
        self.min_delta = min_delta
        self.patience = patience
        self.percentage = percentage
        self.best = None
        self.num_bad_epochs = 0
        self.is_better = self._always if patience == 0 else self._improves

    @staticmethod
    def _always(value, best):
        return True

    def _improves(self, value, best):
        margin = best * self.min_delta / 100 if self.percentage else self.min_delta
        return value > best + margin if self.mode == 'max' else value < best - margin

    def curr_is_better(self, metrics):
        return self.is_better(metrics, self.best)

    def step(self, metrics):
        """Feed one epoch's metric; True = stop now."""
        if self.patience == 0:
            return False
        if self.best is None:
            self.best = metrics
            return False
        if math.isnan(metrics):
            return True
        if self.is_better(metrics, self.best):
            self.best = metrics
            self.num_bad_epochs = 0
        else:
            self.num_bad_epochs += 1
        return self.num_bad_epochs >= self.patience
