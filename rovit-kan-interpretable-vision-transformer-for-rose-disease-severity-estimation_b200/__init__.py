"""B200 (sm_100a) implementation of the RoViT-KAN forward/backward path.

Layout: `csrc/` CUDA kernels + the C ABI (`include/rovitkan.h`), `_lib.py` the ctypes binding,
`ops.py` autograd glue, `models/` and `training/` the host-side mirror of the reference's
`models/{rovit_kan,kan,backbone,heads}.py` and `training/losses.py` (same class names, constructor
signatures, attribute names and state_dict keys), `data/` a stand-in for the reference's git-ignored
`data` package, `dropin.py` / `launch.py` the import hook that lets the reference's scripts run unmodified on
top of all this (`rovitkan_b200.install()`, `python -m rovitkan_b200.launch <script>`).
"""

__version__ = '0.2.0'

from .dropin import install, uninstall  # noqa: E402,F401
