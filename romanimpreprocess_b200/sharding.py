"""Work partitioning across the GPUs of one box.

The reference's only parallelism is one independent OS process per SCA (Slurm array 1-18,
reference runs/summer2025run/OpenUniverse_to_L1L2.job:4): an (exposure, SCA) pair needs its own L1 cube and that SCA's
CALDIR and nothing else.  Here: one process per GPU, work items dealt out statically, NO collective on the data path.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only to agree on timings / counts at the end.
"""


def assign_items(items, rank, world):
    """Items (e.g. ``(exposure, sca)`` pairs) of this rank.

    SCA-major dealing: all exposures of one SCA go to the same rank whenever there are at least ``world`` SCAs, so each
    GPU keeps as few CALDIRs resident as possible (a CALDIR is ~1-2 GB, an exposure 0.27 GB); ranks differ by at most
    one SCA.  With fewer SCAs than ranks the exposures of an SCA are spread over several ranks.
    """
    items = list(items)
    scas = sorted({sca for _, sca in items})
    if len(scas) >= world:
        mine = set(scas[rank::world])
        return [it for it in items if it[1] in mine]
    return items[rank::world]


def assign_items_balanced(items, rank, world):
    """Like ``assign_items``, but balanced when the SCA count is not a multiple of ``world`` (18 SCAs on 8 GPUs: 3/3/2/2/..
    SCA-major leaves six GPUs idle a quarter of the time).  The first ``world * (nsca // world)`` SCAs are dealt
    SCA-major (all their exposures on one rank); the exposures of the leftover SCAs are dealt round-robin over all ranks,
    which costs each rank up to ``nsca % world`` more resident CALDIRs (a few GB of 180) and evens the item counts to
    within one."""
    items = list(items)
    scas = sorted({sca for _, sca in items})
    if len(scas) < world:
        return items[rank::world]
    nfull = world * (len(scas) // world)
    major = set(scas[:nfull][rank::world])
    left = [it for it in items if it[1] in set(scas[nfull:])]
    return [it for it in items if it[1] in major] + left[rank::world]


def imbalance(items, world, assign=None):
    """max / mean of the per-rank item counts (1.0 = perfectly balanced) for a given assignment function."""
    assign = assign or assign_items_balanced
    counts = [len(assign(items, r, world)) for r in range(world)]
    return max(counts) / (sum(counts) / float(world)) if sum(counts) else 1.0


def resident_scas(items):
    """The SCAs whose CALDIR a rank must hold for its items."""
    return sorted({sca for _, sca in items})


def job_throughput(units_this_rank, seconds_this_rank, group=None):
    """Whole-job rate = units processed by all ranks / the slowest rank's time (max over ranks).

    Returns ``(rate, total_units, max_seconds)``.  Works without an initialised process group (single process).
    """
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return units_this_rank / seconds_this_rank, units_this_rank, seconds_this_rank
    if not (dist.is_available() and dist.is_initialized()):
        return units_this_rank / seconds_this_rank, units_this_rank, seconds_this_rank
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(seconds_this_rank)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units_this_rank)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(u.item() / t.item()), float(u.item()), float(t.item())
