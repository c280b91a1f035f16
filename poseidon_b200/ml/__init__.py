"""Generic ML-application helpers of the PMLS tree (reference: ps/src/ml/ — feature vectors, LibSVM / binary data
loading, numerically safe math, per-worker workload partitioning, meta files, a background disk streamer).  The deep
learning application does not use them; they are here so that code written against ``ml/include/ml.hpp`` has a home.

Tensors instead of ``std::vector``: a batch of sparse samples is one CSR triple, a dense one a 2-D tensor."""
from .data_loading import (read_data_label_binary, read_data_label_libsvm,              # noqa: F401
                           read_data_label_sparse_feature_binary, write_sparse_feature_binary)
from .disk_stream import DiskStreamer                                                    # noqa: F401
from .features import DenseFeature, SparseBatch, SparseFeature                           # noqa: F401
from .math_util import (feature_scale_and_add, log_sum, log_sum_vec, safe_log, sigmoid, softmax,   # noqa: F401
                        dot)
from .metafile import MetafileReader                                                     # noqa: F401
from .workload import WorkloadManager, WorkloadManagerConfig                             # noqa: F401
