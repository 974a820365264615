"""Parsers for the structured ``--name`` and ``-j`` component arguments (reference
torchx/components/structured_arg.py:36-152 StructuredNameArgument, :156-236 StructuredJArgument)."""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

from torchx_b200 import specs


@dataclass
class StructuredNameArgument:
    """``{experiment}/{run}``; either side may be empty, a name without ``/`` is the run name.  An empty run name
    falls back to the last module component or the script stem; an empty experiment to ``default-experiment``."""

    experiment_name: str
    run_name: str

    def __str__(self) -> str:
        return f"{self.experiment_name or ''}/{self.run_name}"

    @staticmethod
    def parse_from(name: str, m: Optional[str] = None, script: Optional[str] = None,
                   default_experiment_name: str = "default-experiment") -> "StructuredNameArgument":
        if not m and not script:
            raise ValueError("No main module or script specified. Specify either a main module or a script path")
        if m and script:
            raise ValueError("Both main module and script set. Specify exactly one of: main module or script, but not both")
        experiment, sep, run = name.partition("/")
        if not sep:  # no delimiter: the whole thing is the run name
            experiment, run = "", name
        if not run:
            run = m.rpartition(".")[2] if m else Path(script).stem  # type: ignore[arg-type]
        return StructuredNameArgument(experiment or default_experiment_name, run)


@dataclass
class StructuredJArgument:
    """``{nnodes}[x{nproc_per_node}]`` resolved against a named resource: omitted nproc = the host's GPU count."""

    nnodes: int
    nproc_per_node: int

    def __str__(self) -> str:
        return f"{self.nnodes}x{self.nproc_per_node}"

    @staticmethod
    def parse_from(h: str, j: str) -> "StructuredJArgument":
        parts = j.split("x")
        gpus = specs.named_resources[h].gpu
        if len(parts) == 1:
            nnodes = int(parts[0])
            if gpus <= 0:
                raise ValueError(
                    f"nproc_per_node cannot be inferred from GPU count. `{h}` is not a GPU instance."
                    f" You must specify `-j $NNODESx$NPROCS_PER_NODE` (e.g. `-j {nnodes}x8`)")
            return StructuredJArgument(nnodes, gpus)
        if len(parts) == 2:
            nnodes, nproc = int(parts[0]), int(parts[1])
            if nproc != gpus:
                warnings.warn(
                    f"In `-j {j}` you specified nproc_per_node={nproc} which does not equal the number of GPUs on a {h}: {gpus}."
                    f" This may lead to under-utilization or an error. If this was intentional, ignore this warning."
                    f" Otherwise set `-j {nnodes}` to auto-set nproc_per_node to the number of GPUs on the host.")
            return StructuredJArgument(nnodes, nproc)
        raise ValueError(f"Invalid format for `-j $NNODESx$NPROCS_PER_NODE` (e.g. `-j 1x8`). Given: {j}")
