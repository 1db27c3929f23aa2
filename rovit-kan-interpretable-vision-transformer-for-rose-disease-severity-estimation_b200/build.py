"""Build librovitkan.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python build.py            # incremental: only stale objects are recompiled
    python build.py --force

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the tree to the GPU box.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'librovitkan.so')
SOURCES = ['tma_host.cu', 'gemm.cu', 'attention.cu', 'attention_tc.cu', 'encoder_kernels.cu', 'kan.cu', 'heads.cu', 'encoder.cu', 'optimizer.cu', 'api.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _headers_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    paths.append(os.path.join(HERE, '..', 'include', 'rovitkan.h'))
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_mtime()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace('.cu', '.o'))
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC] + FLAGS + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {s}:\n{r.stdout}\n{r.stderr}')
        return s

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s in ex.map(compile_one, jobs):
                if verbose:
                    print('compiled', os.path.basename(s), flush=True)
    objs = [os.path.join(OBJ, s.replace('.cu', '.o')) for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print('linked', LIB, flush=True)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
