"""Multi-GPU host logic: independent AMPS carriers, one (or more) per GPU, no collective on the
sample path (BASELINE.json north_star).  torch.distributed is used only for the barrier,
the max-over-ranks timing and for gathering the (tiny) per-carrier results on rank 0.
"""
from __future__ import annotations

from dataclasses import dataclass

CHANNEL_SPACING_HZ = 30e3          # AMPS channel raster (870 + 0.03 N MHz)
BASE_OFFSET_HZ = -160e3            # rx_offset of the reference graph (grc/ampsbs.grc:212-238)


@dataclass(frozen=True)
class Carrier:
    index: int
    center_freq: float
    min10: str
    seed: int


def carrier(g: int) -> Carrier:
    """BASELINE config 4: carrier g at -160 kHz + 30 kHz * g, seed 0xA3B5 + g, its own MIN."""
    return Carrier(g, BASE_OFFSET_HZ + CHANNEL_SPACING_HZ * g, "21255512%02d" % (30 + g % 70), 0xA3B5 + g)


def carrier_plan(world: int, n_carriers: int | None = None) -> list[list[Carrier]]:
    """Round-robin assignment of carriers to ranks (one per GPU when n_carriers == world)."""
    n = world if n_carriers is None else n_carriers
    plan: list[list[Carrier]] = [[] for _ in range(world)]
    for g in range(n):
        plan[g % world].append(carrier(g))
    return plan


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a python float (device timings); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_throughput(samples_local: float, seconds_local: float, device=None) -> tuple[float, float]:
    """(total samples over all ranks, max seconds over ranks) -> the bench's whole-job value is their ratio."""
    return sum_over_ranks(samples_local, device), max_over_ranks(seconds_local, device)


def gather_results(local: list, dst: int = 0) -> list | None:
    """Collect per-carrier result objects (decoded MINs, burst counts: bytes per burst) on rank dst."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    return out
