"""Tie-aware k-recall@k exactly as the reference's drivers compute it.

Follows calculate_recall (BANG_Base/test_driver.cpp:43-93, BANG_Inmemory/main.cu:108-163): the ground
truth set for a query is GT[0..t) where t extends past k over entries whose distance equals
GT[k-1]; recall = |GT set ∩ result[0..k)| summed over queries, reported as a percentage
(`total / nq * 100 / k`).
"""
from __future__ import annotations

import numpy as np


def calculate_recall(gt_ids: np.ndarray, gt_dists: np.ndarray | None, results: np.ndarray, recall_at: int) -> float:
    nq = results.shape[0]
    dim_gs = gt_ids.shape[1]
    total = 0
    for i in range(nq):
        t = recall_at
        if gt_dists is not None:
            t = recall_at - 1
            while t < dim_gs and gt_dists[i, t] == gt_dists[i, recall_at - 1]:
                t += 1
        gt = set(int(v) for v in gt_ids[i, :t])
        res = set(int(v) for v in results[i, :recall_at])
        total += len(gt & res)
    return total / nq * (100.0 / recall_at)


def report_header(k: int) -> str:
    # test_driver.cpp:402-403
    return f"L\tTime \tQPS\t\t{k}-r@{k}\n--\t---- \t---\t\t------"


def report_line(L: int, ms: float, qps: float, recall: float) -> str:
    # test_driver.cpp:526 (fixed, precision 2)
    return f"{L}\t{ms:.2f}\t{qps:.2f}\t{recall:.2f}"
