"""Run one of the reference's scripts, unmodified, on top of the sm_100a modules:

    python -m rovitkan_b200.launch /path/to/reference/scripts/train.py --data_root ... --output_dir ...

Installs the import hook (dropin.py), then executes the script as `__main__` exactly as `python script.py`
would (script directory at sys.path[0], sys.argv rewritten).  `--synthetic-data` before the script path forces the
synthetic `data` package even if a real one exists.
"""

from __future__ import annotations

import os
import runpy
import sys

from .dropin import install


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    synthetic = None
    while argv and argv[0].startswith('--'):
        flag = argv.pop(0)
        if flag == '--synthetic-data':
            synthetic = True
        elif flag == '--real-data':
            synthetic = False
        else:
            raise SystemExit(f'rovitkan_b200.launch: unknown option {flag}')
    if not argv:
        raise SystemExit('usage: python -m rovitkan_b200.launch [--synthetic-data|--real-data] <reference script.py> [script args...]')
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f'rovitkan_b200.launch: no such script: {script}')
    install(synthetic_data=synthetic)
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
