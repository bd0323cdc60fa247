from dataclasses import dataclass, field
from typing import Any


@dataclass
class TaskPrediction:
    rank_level: Any
    temperature: float
    predictions: list = field(default_factory=list)  # [(taxon_id, probability), ...] best first


@dataclass
class HierarchicalClassificationResult:
    taxonomy_context: Any
    tasks: list
    subtree_roots: Any = None
