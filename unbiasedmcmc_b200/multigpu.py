"""Multi-GPU host logic: replicates shard by global index, the only collective is a sum all-reduce of the flat
summary buffer (include/ubm.h: ubm_summary_len).  One process per GPU (torchrun); NCCL on GPUs, gloo in CPU tests."""
import numpy as np

from . import _abi as A


def shard_offset(step, rank, world, nrep_per_rank):
    """First global replicate index of `rank` in batch `step` (weak scaling: every rank runs nrep_per_rank)."""
    return (step * world + rank) * nrep_per_rank


def split_replicates(nrep_total, world):
    """Contiguous split of a fixed total (strong scaling): list of (offset, count) per rank."""
    base, rem = divmod(nrep_total, world)
    out, off = [], 0
    for r in range(world):
        c = base + (1 if r < rem else 0)
        out.append((off, c))
        off += c
    return out


def allreduce_summary(summary, group=None):
    """In-place sum all-reduce of a summary buffer: a torch tensor (CUDA for NCCL, CPU for gloo) or a numpy array."""
    import torch
    import torch.distributed as dist
    if isinstance(summary, np.ndarray):
        t = torch.from_numpy(summary)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return summary
    dist.all_reduce(summary, op=dist.ReduceOp.SUM, group=group)
    return summary


def summary_stats(summary, dimh, nbins):
    """What the R user computes from the replicates (README.Rmd:119-121, R/histogram_c_chains.R:40-42):
    mean and standard error of H, of each bin proportion, plus mean meeting time and mean cost."""
    s = np.asarray(summary, dtype=np.float64)
    n = s[0]
    h = A.UBM_SUMMARY_HEAD

    def mean_se(sum1, sum2):
        mean = sum1 / n
        var = np.maximum(sum2 / n - mean * mean, 0.0) * n / max(n - 1.0, 1.0)
        return mean, np.sqrt(var / n)

    hm, hse = mean_se(s[h:h + dimh], s[h + dimh:h + 2 * dimh])
    bm, bse = mean_se(s[h + 2 * dimh:h + 2 * dimh + nbins], s[h + 2 * dimh + nbins:h + 2 * dimh + 2 * nbins])
    return {"n_done": n, "n_censored": s[1], "mean_meetingtime": s[2] / n, "mean_cost": s[4] / n,
            "h_mean": hm, "h_se": hse, "bin_mean": bm, "bin_se": bse}
