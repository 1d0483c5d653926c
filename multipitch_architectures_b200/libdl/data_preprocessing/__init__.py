from .hcqt import (compute_hopsize_cqt, compute_hcqt, compute_efficient_hcqt, compute_annotation_array_nooverlap, estimate_tuning,
                   HCQTPlan, get_plan)
