"""nann_b200 -- B200-native implementation of alibaba/nann's model-scored HNSW retrieval hot path.

Importing the package loads libnann_b200.so (built in-tree by `python -m nann_b200.build`);
it fails loudly when the library is missing.  There is no CPU or PyTorch fallback.
"""
from . import _lib

_lib.lib()  # raise ImportError now rather than at first use

from ._lib import NannError, SCORER_EXACT, SCORER_TENSOR, launch_count, device_info  # noqa: E402
from . import ops  # noqa: E402
from .ops import (group_gather, bitmap_ref_difference, bloom_filter_difference, top_k, batch_top_k_on_rt, gather, huge_const, HugeConst,  # noqa: E402
                  Scorer, blaze_xla_op, score_ids, merge_topk)
from .retrieval import Index, Searcher, EvalSearcher, retrieve_opwise  # noqa: E402

__all__ = ["NannError", "SCORER_EXACT", "SCORER_TENSOR", "launch_count", "device_info", "ops",
           "group_gather", "bitmap_ref_difference", "bloom_filter_difference", "top_k", "batch_top_k_on_rt", "gather", "huge_const", "HugeConst",
           "Scorer", "blaze_xla_op", "score_ids", "merge_topk", "Index", "Searcher", "EvalSearcher", "retrieve_opwise"]
