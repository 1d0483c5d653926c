"""Mirror of the reference's `libdl.metrics` surface: evaluation measures (one float64 pass on the device, see eval_metrics.py) and the
early-stopping helper of the training scripts (host control flow, see monitoring.py)."""
from . import eval_metrics as _em
from . import monitoring as _mon

early_stopping = _mon.early_stopping
calculate_single_measure = _em.calculate_single_measure
calculate_eval_measures = _em.calculate_eval_measures
calculate_mpe_measures_mireval = _em.calculate_mpe_measures_mireval
# additions of this package (not in the reference): the libfmp-shaped counting entry point, the raw device sums, rank statistics
compute_eval_measures = _em.compute_eval_measures
eval_sums = _em.eval_sums
roc_auc = _em.roc_auc
average_precision = _em.average_precision

__all__ = ['early_stopping', 'calculate_single_measure', 'calculate_eval_measures', 'calculate_mpe_measures_mireval',
           'compute_eval_measures', 'eval_sums', 'roc_auc', 'average_precision']
