"""One HSL layer per GPU: the host-side choreography of eQ's "layer parallel" MPI model.

Upstream, rank 0 is the controller (cells) and every HSL species lives on its own QS rank
(src/simulation.cpp:25,38-42,628-671); per step the controller exchanges whole fields with each QS rank
(src/simulation.cpp:428-432,455-463,491-505).  Here every rank of a torch.distributed job owns the layers
``l % world == rank`` on its GPU, the fields never leave HBM, and only per-cell vectors move: each rank
gathers (readHSL) for its layers, the per-layer samples are all-gathered so that the gene-circuit model
(Strain::computeProteins, untouched, CPU) can couple the species, and the resulting per-layer deposits are
scattered (writeHSL) where the layer lives.  There is no data-path collective on the field itself.

The solver objects are injected (``solver_factory``) so the choreography can be exercised on CPU with the
gloo backend; the product factory is ``eq_b200.GpuHSL``.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def layer_owner(layer: int, world: int) -> int:
    """Rank that owns a layer.  Mirrors the rank->layer map of mpiAssignCommunicators
    (src/simulation.cpp:628-645) without the dedicated controller rank."""
    return layer % world


def check_layout(num_layers: int, world: int) -> None:
    """The reference insists on (npes-1) % numHSLGrids == 0 (src/simulation.cpp:38-42); the GPU layout only
    needs every rank to have work when there are at least as many layers as ranks."""
    if num_layers < 1:
        raise ValueError("need at least one HSL layer")
    if world < 1:
        raise ValueError("world size must be positive")


class LayerGroup:
    def __init__(self, d_hsl: Sequence[float], solver_factory: Callable[[int, float], object],
                 rank: int | None = None, world: int | None = None, device: str = "cpu"):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        check_layout(len(d_hsl), self.world)
        self.num_layers = len(d_hsl)
        self.device = device
        self.local = {l: solver_factory(l, d_hsl[l]) for l in range(self.num_layers)
                      if layer_owner(l, self.world) == self.rank}

    def upload_cells(self, records: np.ndarray, nodes_per_micron: float) -> None:
        """Every layer needs the same rod geometry (20k x 128 B; cf. the 5*N doubles upstream)."""
        self.ncells = len(records)
        for s in self.local.values():
            s.upload_cells(records, nodes_per_micron)

    def gather_all(self) -> np.ndarray:
        """hslData of every cell for every layer, identical on all ranks: [num_layers, ncells]."""
        out = torch.zeros(self.num_layers, self.ncells, dtype=torch.float64, device=self.device)
        for l, s in self.local.items():
            out[l] = torch.from_numpy(np.asarray(s.gather())).to(self.device)
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)  # disjoint rows: a sum is an all-gather
        return out.cpu().numpy()

    def scatter_all(self, amounts: np.ndarray) -> None:
        """amounts[l, k] = Strain's deltaHSL of cell k for layer l (same array on every rank)."""
        for l, s in self.local.items():
            s.scatter(amounts[l])

    def step(self) -> None:
        for s in self.local.values():
            s.step()

    def boundary_flux(self) -> np.ndarray:
        f = torch.zeros(self.num_layers, dtype=torch.float64, device=self.device)
        for l, s in self.local.items():
            f[l] = s.totalBoundaryFlux
        if self.world > 1:
            dist.all_reduce(f, op=dist.ReduceOp.SUM)
        return f.cpu().numpy()


def max_over_ranks_ms(ms: float, device: str = "cpu") -> float:
    """Timing rule: a multi-GPU number is the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
