"""Lazy access to the autograd glue (ops.py) for the module mirrors: importing `rovitkan_b200.models` must not
load torch.amp / ctypes machinery until a forward actually runs."""

import importlib


def ops():
    return importlib.import_module(__package__ + '.ops')
