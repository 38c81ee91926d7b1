"""Type aliases of the per-pair solver outputs (reference ``src/objects/emdu_out.py:7-9``)."""
from typing import Dict, NewType, Tuple

UnifracDistance = NewType('UnifracDistance', float)
EmdFlow = NewType('EmdFlow', Dict[Tuple[int, int], float])
DifferentialAbundance = NewType('DifferentialAbundance', Dict[Tuple[int, int], float])
