"""Names of the files the all-pairs drivers write; ``{0}`` is the run identifier (``-i``).  The names are part of the
drop-in contract: they equal the templates of the reference's ``src/utility/constant.py:4-8``."""
_STEM = "fununifrac_out_{0}"

FUNUNIFRAC_OUT__MAIN_FILE_NAME = _STEM + ".main.npy"                    # float64 N x N distance matrix (np.save)
FUNUNIFRAC_OUT__MAIN_BASIS__FILE_NAME = _STEM + ".main.basis.npy"       # text: one sample path per line, matrix order
FUNUNIFRAC_OUT__DIFFAB_FILE_NAME = _STEM + ".diffab"                    # sparse.save_npz appends .npz
FUNUNIFRAC_OUT__DIFFAB_NODES_FILE_NAME = _STEM + ".diffab.nodes.npy"    # text: node names in basis order
FUNUNIFRAC_OUT__PUSH_FILE_NAME = _STEM + ".push.pkl"                    # pickle: {sample path: pushed vector}
