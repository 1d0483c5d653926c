from .hcqt_datasets import dataset_context
