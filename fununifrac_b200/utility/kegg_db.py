"""BRITE id whitelist used only to validate ``-b`` (reference ``src/utility/kegg_db.py:19-28``,
consumed at ``src/compute_all_pairwise_fununifrac.py:37-38``)."""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_BRITES_FILE = os.path.join(os.path.dirname(_HERE), "data", "brites.txt")


class _KeggDatabase:
    def __init__(self) -> None:
        self._brites = None

    @property
    def brites(self):
        if self._brites is None:
            with open(_BRITES_FILE, 'r') as f:
                self._brites = [str(l).strip() for l in f.readlines()]
        return self._brites


# singleton is enough
instance = _KeggDatabase()
