"""fununifrac_b200 — B200-native all-pairs functional UniFrac.

Drop-in for the ``compute_fununifrac.py`` → ``src/compute_all_pairwise_fununifrac.py``
path of KoslickiLab/FunUniFrac.  Host code is Python; all arithmetic of the
path runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/fununifrac.h`` (``fununifrac_b200/csrc``).  There is no CPU fallback:
using the solver without the built extension or without a GPU raises.
"""
__version__ = "0.1.0"
