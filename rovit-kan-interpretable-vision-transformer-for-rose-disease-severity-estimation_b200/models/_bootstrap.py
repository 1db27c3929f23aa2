"""Resolve the product package whether this directory is imported as `rovitkan_b200.models`
(normal use) or as the top-level package `models` (drop-in for the reference's scripts, which do
`from models.rovit_kan import RoViTKAN`)."""

import importlib
import os
import sys


def product_package():
    try:
        return importlib.import_module('rovitkan_b200')
    except ImportError:
        repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        if repo_root not in sys.path:
            sys.path.append(repo_root)
        return importlib.import_module('rovitkan_b200')


def ops():
    product_package()
    return importlib.import_module('rovitkan_b200.ops')
