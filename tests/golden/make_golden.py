#!/usr/bin/env python
"""Writes tests/golden/kat_ops.json: the known-answer vectors the reference's own tests hold for
the hot path.  The reference cannot be executed here (patched TF 1.15 + bazel), so every expected
value below is either (a) asserted by a reference test file, or (b) the value stated in a
reference test's docstring/comment, or (c) hand-derived from the cited kernel source for the
exact inputs a reference print-only test feeds.  Each record says which.

Paths are relative to the reference checkout; UO = tensorflow/tensorflow/core/user_ops.
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

kat = {
    "group_gather": [
        {   # UO/beam_search_op/group_gather_test.py:18-26 (inputs); expected: (c) from GroupGather_kernel.cc:136-170
            "source": "UO/beam_search_op/group_gather_test.py:18-26",
            "kind": "hand-derived from kernel source for the test's inputs",
            "dtype": "int64",
            "params_values": [0, 1, 1, 2, 3, 4, 3, 4, 5, 5, 6, 7, 8, 8, 9, 10, 11, 12],
            "params_row_splits": [0, 6, 11, 15, 18],
            "indices_values": [0, 1, 3], "indices_row_splits": [0, 2, 3],
            "unique": False,
            "ret_values": [0, 1, 1, 2, 3, 4, 3, 4, 5, 5, 6, 10, 11, 12], "ret_row_splits": [0, 11, 14],
        },
        {   # docstring example group_gather_test.py:7-10: t=[[0,1],[2,3,4],[5,6],[7,8,9]], ind=[[0,1],[3]] -> [[0,1,2,3,4],[7,8,9]]
            "source": "UO/beam_search_op/group_gather_test.py:7-10 (docstring)",
            "kind": "stated in reference docstring",
            "dtype": "int64",
            "params_values": [0, 1, 2, 3, 4, 5, 6, 7, 8, 9], "params_row_splits": [0, 2, 5, 7, 10],
            "indices_values": [0, 1, 3], "indices_row_splits": [0, 2, 3],
            "unique": False,
            "ret_values": [0, 1, 2, 3, 4, 7, 8, 9], "ret_row_splits": [0, 5, 8],
        },
        {   # same call with unique=True: the reference's order is unordered_set order; the SETS are pinned
            "source": "UO/beam_search_op/group_gather_test.py:23",
            "kind": "hand-derived (set semantics, GroupGather_kernel.cc:91-131)",
            "dtype": "int64",
            "params_values": [0, 1, 1, 2, 3, 4, 3, 4, 5, 5, 6, 7, 8, 8, 9, 10, 11, 12],
            "params_row_splits": [0, 6, 11, 15, 18],
            "indices_values": [0, 1, 3], "indices_row_splits": [0, 2, 3],
            "unique": True,
            "ret_sets": [[0, 1, 2, 3, 4, 5, 6], [10, 11, 12]], "ret_row_splits": [0, 7, 10],
        },
        {   # empty params (group_gather_test.py:20,24): void -> values=[], row_splits=[0]  (GroupGather_kernel.cc:69-77)
            "source": "UO/beam_search_op/group_gather_test.py:24",
            "kind": "hand-derived (void-input branch)",
            "dtype": "int64",
            "params_values": [], "params_row_splits": [0],
            "indices_values": [0, 1, 3], "indices_row_splits": [0, 2, 3],
            "unique": False, "ret_values": [], "ret_row_splits": [0],
        },
        {   # empty indices (group_gather_test.py:25)
            "source": "UO/beam_search_op/group_gather_test.py:25",
            "kind": "hand-derived (void-input branch)",
            "dtype": "int64",
            "params_values": [0, 1, 1, 2, 3, 4, 3, 4, 5, 5, 6, 7, 8, 8, 9, 10, 11, 12],
            "params_row_splits": [0, 6, 11, 15, 18],
            "indices_values": [], "indices_row_splits": [0],
            "unique": False, "ret_values": [], "ret_row_splits": [0],
        },
    ],
    # UO/bitmap_op/bitmap_ref_difference.py:16-29: three chained calls on ONE flags variable
    "bitmap_ref_difference_chain": {
        "source": "UO/bitmap_op/bitmap_ref_difference.py:16-29",
        "kind": "hand-derived from bitmap_ops.cc:221-234 for the test's inputs",
        "dtype": "int32",
        "flags0": [0, 0, 0, 0],
        "calls": [
            {"values": [1, 1, 2, 2, 3, 4, 5, 11, 12, 13], "row_splits": [0, 7, 10],
             "c_values": [1, 2, 3, 4, 5, 11, 12, 13], "c_row_splits": [0, 5, 8]},
            {"values": [4, 5, 6, 7, 7, 8, 10, 13, 14], "row_splits": [0, 7, 9],
             "c_values": [6, 7, 8, 10, 14], "c_row_splits": [0, 4, 5]},
            {"values": [4, 5, 6, 7, 7, 8, 10, 13, 14], "row_splits": [0, 7, 9],
             "c_values": [], "c_row_splits": [0, 0, 0]},
        ],
        "flags_final": [32254, 0, 0, 0],   # bits 1-8, 10-14
    },
    # tensorflow/tensorflow/python/kernel_tests/topk_op_test.py -- ASSERTED by the reference test-suite
    "topk_v2": [
        {"source": "topk_op_test.py:97-99 testTop1", "kind": "asserted by reference test",
         "input": [[0.1, 0.3, 0.2, 0.4], [0.1, 0.3, 0.3, 0.2]], "k": 1,
         "values": [[0.4], [0.3]], "indices": [[3], [1]]},
        {"source": "topk_op_test.py:101-103 testTop2", "kind": "asserted by reference test",
         "input": [[0.1, 0.3, 0.2, 0.4], [0.1, 0.3, 0.4, 0.2]], "k": 2,
         "values": [[0.4, 0.3], [0.4, 0.3]], "indices": [[3, 1], [2, 1]]},
        {"source": "topk_op_test.py:166-169 testTopAll", "kind": "asserted by reference test",
         "input": [[0.1, 0.3, 0.2, 0.4], [0.1, 0.3, 0.3, 0.2]], "k": 4,
         "values": [[0.4, 0.3, 0.2, 0.1], [0.3, 0.3, 0.2, 0.1]], "indices": [[3, 1, 2, 0], [1, 2, 3, 0]]},
        {"source": "topk_op_test.py:178-180 testTop3Vector", "kind": "asserted by reference test",
         "input": [3, 6, 15, 18, 6, 12, 1, 17, 3, 0, 4, 19, 1, 6], "k": 3,
         "values": [19, 18, 17], "indices": [11, 3, 7]},
        {"source": "SURVEY Appendix D (topk_op.cc:142-150 tie rule)", "kind": "hand-derived",
         "input": [1, 3, 3, 2], "k": 2, "values": [3, 3], "indices": [1, 2]},
    ],
    "topk_v2_errors": [
        {"source": "topk_op_test.py:192-199 testKNegative", "input": [[0.1, 0.2], [0.3, 0.4]], "k": -7,
         "message": "Need k >= 0, got -7"},
        {"source": "topk_op_test.py:202-208 testKTooLarge", "input": [[0.1, 0.2], [0.3, 0.4]], "k": 4,
         "message": "input must have at least k columns"},
    ],
    # UO/topk_op/batch_topk_on_rt_test.py:12-17 (print-only; expected values hand-derived from
    # BatchTopKOnRT_kernel.cc:110-148; all inputs distinct so the unspecified tie order does not matter)
    "batch_topk_on_rt": {
        "source": "UO/topk_op/batch_topk_on_rt_test.py:12-17",
        "values": [1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 26, 27, 31, 32, 33, 34, 35],
        "row_splits": [0, 7, 12, 19, 24],
        "calls": [
            {"k": [3, 2, 3, 2], "ascending": False,
             "values_out": [7, 6, 5, 15, 14, 27, 26, 25, 35, 34], "idx_out": [6, 5, 4, 4, 3, 6, 5, 4, 4, 3],
             "row_splits_out": [0, 3, 5, 8, 10]},
            {"k": 6, "ascending": True,
             "values_out": [1, 2, 3, 4, 5, 6, 11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 26, 31, 32, 33, 34, 35],
             "idx_out": [0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4],
             "row_splits_out": [0, 6, 11, 17, 22]},
        ],
        "void": {"values": [], "row_splits": [0], "k": 3, "values_out": [], "idx_out": [], "row_splits_out": [0]},
    },
    # "should be" comments of the reference's own scripts
    "batch_gather_on_rt": {"source": "UO/beam_search_op/batch_gather_on_rt_test.py:17-35", "kind": "stated in reference comments",
        "params_values": [1, 2, 3, 4, 5], "params_row_splits": [0, 3, 5], "indices_values": [0, 1, 1],
        "indices_row_splits": [0, 2, 3], "ret_values": [1, 2, 5], "ret_row_splits": [0, 2, 3]},
    "batch_concat_on_rt": {"source": "UO/beam_search_op/batch_concat_on_rt_test.py:13-35", "kind": "stated in reference comments",
        "left_values": [1, 2, 3, 4, 5], "left_row_splits": [0, 3, 5], "right_values": [0, 1, 1], "right_row_splits": [0, 2, 3],
        "ret_values": [1, 2, 3, 0, 1, 4, 5, 1], "ret_row_splits": [0, 5, 8]},
    "splits_gather": {"source": "UO/beam_search_op/splits_gather_test.py:7-10,18-24", "kind": "stated in reference docstring",
        "splits": [0, 2, 5, 7, 10], "indices_values": [0, 1, 3], "indices_row_splits": [0, 2, 3],
        "ret_values": [0, 1, 2, 3, 4, 7, 8, 9], "ret_row_splits": [0, 5, 8]},
    # farmhash::Fingerprint64 values the reference's own tests state (BloomFilterDifference hashes std::to_string(node) with it)
    "fingerprint64": [
        {"source": "tensorflow/python/kernel_tests/string_to_hash_bucket_op_test.py:46-49", "kind": "stated in reference test comments (the mod-10 buckets are asserted)",
         "inputs": ["a", "b", "c", "d"], "values": [12917804110809363939, 11795596070477164822, 11430444447143000872, 4470636696479570465],
         "mod10": [9, 2, 2, 5]},
        {"source": "tensorflow/core/platform/fingerprint_test.cc:27-28", "kind": "asserted by reference test",
         "inputs": ["Hello", "World"], "values": [15404698994557526151, 18308117990299812472]},
    ],
    # UO/bitmap_op/bloom_filter_difference.py:8-31: three chained calls on ONE flags variable (bucket=0, bucket_size=10);
    # the script only prints, so the expected outputs are hand-derived from bitmap_ops.cc:334-359 (none of these 17 values
    # collides in the 320-bit filter, so the results equal the exact-bitmap ones) and the final flags from the hash chain
    "bloom_filter_difference_chain": {
        "source": "UO/bitmap_op/bloom_filter_difference.py:8-31",
        "kind": "hand-derived from bitmap_ops.cc:334-359 for the test's inputs",
        "dtype": "int32", "bucket": 0, "bucket_size": 10, "flags0": [0] * 10,
        "primes": [9277, 15031, 21433, 26557],           # find_prime_lower_than(29|47|67|83 * 10 * 32), :404-421
        "calls": [
            {"values": [1, 1, 2, 2, 3, 4, 5, 11, 12, 13], "row_splits": [0, 7, 10],
             "c_values": [1, 2, 3, 4, 5, 11, 12, 13], "c_row_splits": [0, 5, 8]},
            {"values": [4, 5, 6, 7, 7, 8, 10, 1000, 13, 14], "row_splits": [0, 7, 10],
             "c_values": [6, 7, 8, 10, 1000, 14], "c_row_splits": [0, 4, 6]},
            {"values": [4, 5, 6, 7, 7, 8, 10, 1000, 13, 14], "row_splits": [0, 7, 10],
             "c_values": [], "c_row_splits": [0, 0, 0]},
        ],
        "flags_final": [536872960, 557842448, 541098520, 8194, -1598683056, -918413312, 176161152, -1560247680, 1090717705, 285212705],
    },
    # UO/huge_const_op/huge_const_test.py:6-27 -- arrays saved then read back through HugeConst
    "huge_const": [
        {"source": "UO/huge_const_op/huge_const_test.py:6,22,25", "dtype": "int32", "array": [[1, 2], [3, 4], [5, 6]]},
        {"source": "UO/huge_const_op/huge_const_test.py:8,23,26", "dtype": "int64", "array": [[1, 2, 3], [4, 5, 6]]},
        {"source": "UO/huge_const_op/huge_const_test.py:10,24,27", "dtype": "float32", "array": [[1, 2, 3, 4, 5, 6]]},
    ],
}

with open(os.path.join(HERE, "kat_ops.json"), "w") as f:
    json.dump(kat, f, indent=1)
print("wrote", os.path.join(HERE, "kat_ops.json"))
