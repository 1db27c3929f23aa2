"""Import hook that puts this implementation under the module names the reference's scripts import.

The reference's entry points (`scripts/train.py:9-19`, `scripts/evaluate.py:7-13`, `training/trainer.py:9`,
`evaluation/evaluator.py:230`) insert their own project root at `sys.path[0]` and then import
`models.rovit_kan`, `training.losses`, `training.optimizer`, `training.trainer`, `data.dataset`,
`data.transforms`, `configs.config`, `evaluation.*`, `results.*`.  A `sys.path` entry of ours can never win
against that insert, so `install()` registers a `sys.meta_path` finder (meta-path finders run before the
path-based finder) that serves

    models, models.*      -> rovitkan_b200.models(.*)          (the sm_100a modules)
    training.losses       -> rovitkan_b200.training.losses     (the fused joint loss)
    training              -> an empty package whose __path__ is the reference's own `training/` directory,
                             so `training.trainer` / `training.optimizer` keep coming from the reference
    data, data.*          -> rovitkan_b200.data(.*)            ONLY when no real `data/dataset.py` is importable
                             (the reference git-ignores its `data/` package, .gitignore:60)

and lets every other name fall through to the reference tree.  The aliases are the very same module objects
as `rovitkan_b200.*` (no second copy of the classes).
"""

from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

_PKG = 'rovitkan_b200'
_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_finder = None


def _is_ours(path: str) -> bool:
    try:
        return os.path.commonpath([os.path.abspath(path), _PKG_DIR]) == _PKG_DIR
    except ValueError:
        return False


def _reference_dirs(package: str):
    """Directories `<entry>/<package>` of the CURRENT sys.path that are not ours (the reference's own package)."""
    seen, out = set(), []
    for entry in sys.path:
        d = os.path.join(entry or os.getcwd(), package)
        if d in seen or not os.path.isdir(d) or _is_ours(d):
            continue
        seen.add(d)
        out.append(d)
    return out


def real_data_package() -> str | None:
    """Path of an importable non-synthetic `data` package (one that has dataset.py), if any."""
    for d in _reference_dirs('data'):
        if os.path.exists(os.path.join(d, 'dataset.py')):
            return d
    return None


class _MergedPath:
    """`__path__` of the synthetic `training` package: recomputed from sys.path on every use, because the
    reference's scripts insert their root into sys.path after this hook is installed."""

    def __init__(self, package: str):
        self._package = package

    def _dirs(self):
        return _reference_dirs(self._package)

    def __iter__(self):
        return iter(self._dirs())

    def __len__(self):
        return len(self._dirs())

    def __getitem__(self, i):
        return self._dirs()[i]

    def __contains__(self, item):
        return item in self._dirs()

    def __repr__(self):
        return f'_MergedPath({self._dirs()!r})'


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name: str):
        self.real_name = real_name
        self._spec = None

    def create_module(self, spec):
        module = importlib.import_module(self.real_name)
        self._spec = module.__spec__
        return module

    def exec_module(self, module):
        module.__spec__ = self._spec          # the import system pointed it at the alias spec; undo


class _PackageLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__doc__ = ('merged package installed by rovitkan_b200.install(): `training.losses` is the sm_100a joint loss, every '
                     'other submodule is the reference\'s own file')
        return m

    def exec_module(self, module):
        module.__path__ = _MergedPath(module.__name__)


class DropinFinder(importlib.abc.MetaPathFinder):
    ALIASES = {'models': f'{_PKG}.models', 'training.losses': f'{_PKG}.training.losses', 'data': f'{_PKG}.data'}

    def __init__(self, synthetic_data):
        self.synthetic_data = synthetic_data       # True / False / None (= only if no real data package is importable)
        self.created = []

    def _alias(self, fullname: str) -> str | None:
        top = fullname.split('.', 1)[0]
        if top == 'models':
            return f'{_PKG}.{fullname}'
        if fullname == 'training.losses':
            return self.ALIASES[fullname]
        if top == 'data':
            use = self.synthetic_data
            if use is None:
                use = real_data_package() is None
            return f'{_PKG}.{fullname}' if use else None
        return None

    def find_spec(self, fullname, path=None, target=None):
        if fullname == 'training':
            self.created.append(fullname)
            return importlib.machinery.ModuleSpec(fullname, _PackageLoader(), is_package=True)
        real = self._alias(fullname)
        if real is None:
            return None
        try:
            is_pkg = hasattr(importlib.import_module(real), '__path__')
        except ModuleNotFoundError as e:
            if e.name == real:
                return None
            raise
        self.created.append(fullname)
        return importlib.machinery.ModuleSpec(fullname, _AliasLoader(real), is_package=is_pkg)


def install(synthetic_data: bool | None = None) -> DropinFinder:
    """Register the drop-in finder (idempotent).  `synthetic_data`: True = always serve the synthetic `data`
    package, False = never, None = only when no real `data/dataset.py` is importable."""
    global _finder
    if _finder is not None:
        _finder.synthetic_data = synthetic_data
        return _finder
    clash = [n for n in sys.modules if n.split('.', 1)[0] in ('models', 'training')
             and not getattr(sys.modules[n], '__name__', n).startswith(_PKG)]
    if clash:
        raise RuntimeError(f'rovitkan_b200.install() must run before the reference\'s packages are imported; already loaded: {clash[:4]}')
    _finder = DropinFinder(synthetic_data)
    sys.meta_path.insert(0, _finder)
    return _finder


def uninstall() -> None:
    global _finder
    if _finder is None:
        return
    if _finder in sys.meta_path:
        sys.meta_path.remove(_finder)
    for name in list(sys.modules):
        if name.split('.', 1)[0] in ('models', 'training', 'data') and (name in _finder.created or name.startswith('training.')):
            del sys.modules[name]
    _finder = None
