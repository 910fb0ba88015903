"""sagnn_b200 -- B200-native interval-graph propagation for SelfGNN (LIU-YUXI/SA-GNN).

Only the reference's short-term graph-propagation hot path (model.py:80-92,118-134 and
its backward) lives here: hand-written sm_100a CUDA kernels behind a C ABI
(``include/sagnn_b200.h``, ``lib/libsagnn_b200.so``), a ctypes binding, a
``torch.autograd.Function`` shim and the host-side mirror of the reference's data layer.
There is no CPU fallback: compute entry points raise if the CUDA library or a GPU is missing.
"""
from . import data_handler                                            # noqa: F401
from .data_handler import transToLsts, trans_to_lsts, transpose       # noqa: F401
from .propagate import (Plan, build_plan, bucket_events, propagate, message_propagate, pair_scores,  # noqa: F401
                        propagate_host, host_forward, host_backward, IntervalPropagation)
from .fusion import IntervalFusion, SequenceAttention, SslHead, slabs_to_rtd                       # noqa: F401
from .np_sampler import ReferenceStream                                # noqa: F401
from ._lib import lib_path, load_library, SagnnError                  # noqa: F401

__all__ = [
    "data_handler", "transToLsts", "trans_to_lsts", "transpose", "Plan", "build_plan", "bucket_events", "propagate",
    "message_propagate", "pair_scores", "propagate_host", "host_forward", "host_backward", "IntervalPropagation", "lib_path", "load_library",
    "SagnnError", "IntervalFusion", "SequenceAttention", "SslHead", "slabs_to_rtd", "ReferenceStream",
]
