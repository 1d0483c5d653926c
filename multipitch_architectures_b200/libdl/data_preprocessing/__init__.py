"""Mirror of the reference's `libdl.data_preprocessing` surface: HCQT feature extraction on the GPU (hcqt.py: decimator chain, fused FFT +
constant-Q rows, tuning estimate behind the C ABI) and the host-side annotation rasteriser / hop-size arithmetic."""
from . import hcqt as _hcqt

compute_hopsize_cqt = _hcqt.compute_hopsize_cqt
compute_hcqt = _hcqt.compute_hcqt
compute_efficient_hcqt = _hcqt.compute_efficient_hcqt
compute_annotation_array_nooverlap = _hcqt.compute_annotation_array_nooverlap
# additions of this package: the tuning estimate on its own and the reusable device plan (filter tables + launch schedule)
estimate_tuning = _hcqt.estimate_tuning
HCQTPlan = _hcqt.HCQTPlan
get_plan = _hcqt.get_plan

__all__ = ['compute_hopsize_cqt', 'compute_hcqt', 'compute_efficient_hcqt', 'compute_annotation_array_nooverlap', 'estimate_tuning',
           'HCQTPlan', 'get_plan']
