"""Host-side logic of the multi-GPU path: independent ladders are sharded across ranks, nothing is
exchanged while sampling, and only traces / pooled statistics are gathered afterwards.

On GPUs the gather and the pooled sums go through NCCL inside libaptb200.so (apt_comm_*, apt_run_args.gather).  The
functions here are the rank arithmetic and the statistics pooling (numpy only); tests/test_sharding_gloo.py drives them
over 2 CPU ranks with the gloo backend."""
import numpy as np


def shard_ladders(total_ladders, world_size, rank):
    """Contiguous block partition: returns (ladder0, n_local).  Global ladder ids key the Philox stream
    (include/aptb200.h), so a sharded run reproduces the single-process run ladder by ladder."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_ladders), int(world_size))
    n_local = base + (1 if rank < rem else 0)
    ladder0 = rank * base + min(rank, rem)
    return ladder0, n_local


def local_moments(samples):
    """samples [ladders, rows, cols] -> vector (n, sum_c..., sumsq_c...) that pools by addition."""
    s = np.asarray(samples, dtype=np.float64)
    flat = s.reshape(-1, s.shape[-1])
    return np.concatenate([[flat.shape[0]], flat.sum(0), (flat * flat).sum(0)])


def pooled_mean_var(moments, ncols):
    n = moments[0]
    mean = moments[1:1 + ncols] / n
    var = moments[1 + ncols:1 + 2 * ncols] / n - mean * mean
    return mean, var * n / max(n - 1.0, 1.0)


def ess_batch_means(x):
    """ESS by non-overlapping batch means (sqrt(n) batches); applied identically to GPU and CPU traces
    (coda::effectiveSize of demo/APT_demo.R:203 is not available here)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    b = max(2, int(np.sqrt(n)))
    m = n // b
    if m < 2:
        return float(n)
    bm = x[:b * m].reshape(b, m).mean(1)
    v = x.var(ddof=1)
    vb = bm.var(ddof=1) * m
    return float(n * v / vb) if vb > 0 else float(n)
