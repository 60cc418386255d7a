"""Keypoint sharding of the multi-GPU mode — host-side mirror of what the kernels compute.

Rank r of G owns the contiguous keypoint range [K r // G, K (r + 1) // G) (k_gn_iterate / k_gn_persistent:
`lo = K * shard_rank / shard_world`). Per Gauss-Newton iteration the ranks exchange ONE all-reduce (sum) of the
96-double accumulator: 78 upper-triangle entries of JTJ, 12 of JTr, and the residual / keypoint counters.
"""
ACCUMULATOR_DOUBLES = 96


def shard_bounds(num_keypoints, rank, world):
    return (num_keypoints * rank) // world, (num_keypoints * (rank + 1)) // world


def pack_normal_equations(A, b, n_used):
    """(12x12 A, 12 b, count) → the 96-double accumulator layout (upper triangle row-major, then b, then count)."""
    import numpy as np
    acc = np.zeros(ACCUMULATOR_DOUBLES)
    iu = np.triu_indices(12)
    acc[:78] = np.asarray(A)[iu]
    acc[78:90] = b
    acc[90] = n_used
    return acc


def unpack_normal_equations(acc):
    import numpy as np
    A = np.zeros((12, 12))
    iu = np.triu_indices(12)
    A[iu] = acc[:78]
    A = A + A.T - np.diag(np.diag(A))
    return A, np.array(acc[78:90]), int(round(acc[90]))


def prefix_shares(valid_counts, rank, max_num_residuals):
    """Solvers CERES / ROBUST keep the first `max_num_residuals` valid residual blocks in keypoint order
    (GetProblem, src/ct_icp/ct_icp.cpp:409-424). With contiguous keypoint shards that prefix is split as computed by
    k_lm_select: `valid_counts[r]` = valid blocks on rank r (all-gathered as a sum of one-hot vectors);
    returns (blocks of lower ranks, this rank's share, blocks in the whole problem)."""
    limit = max_num_residuals if max_num_residuals > 0 else 0x7FFFFFFF
    before = int(sum(valid_counts[:rank]))
    total = int(sum(valid_counts))
    share = min(max(limit - before, 0), int(valid_counts[rank]))
    return before, share, min(total, limit)
