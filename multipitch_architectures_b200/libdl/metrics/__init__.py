from .monitoring import early_stopping
from .eval_metrics import (calculate_eval_measures, calculate_single_measure, calculate_mpe_measures_mireval,
                           compute_eval_measures, eval_sums, roc_auc, average_precision)
