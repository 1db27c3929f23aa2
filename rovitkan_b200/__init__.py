"""Importable alias of the product package.

The package directory carries the repository's long, hyphenated name
(`rovit-kan-interpretable-vision-transformer-for-rose-disease-severity-estimation_b200/`), which is
not a Python identifier; this stub points `rovitkan_b200.__path__` at it, so
`import rovitkan_b200.models.kan` resolves to the files inside that directory.
"""

import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         'rovit-kan-interpretable-vision-transformer-for-rose-disease-severity-estimation_b200')
__path__ = [_PKG_DIR]
PACKAGE_DIR = _PKG_DIR

with open(_os.path.join(_PKG_DIR, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, '__init__.py'), 'exec'))
