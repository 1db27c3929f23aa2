"""B200 (sm_100a) implementation of the RoViT-KAN forward/backward path.

Layout: `csrc/` CUDA kernels + the C ABI (`include/rovitkan.h`), `_lib.py` the ctypes binding,
`ops.py` autograd glue, `models/` and `training/` the host-side mirror of the reference's
`models/{rovit_kan,kan,backbone,heads}.py` and `training/losses.py` (same class names, constructor
signatures, attribute names and state_dict keys).  To drop into the reference's scripts, put
`dropin_path()` at the front of `sys.path`.
"""

__version__ = '0.1.0'


def dropin_path() -> str:
    """Directory to put on sys.path so `import models...` / `import training.losses` resolve here."""
    return __path__[0]
