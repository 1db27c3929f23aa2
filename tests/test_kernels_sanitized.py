"""The emulated kernels under AddressSanitizer + UndefinedBehaviorSanitizer (the CPU stand-in for `compute-sanitizer --tool memcheck`,
which needs a GPU): the host-emulation suite (tests/test_kernels_on_host.py) is re-run in a subprocess with the harnesses compiled
`-fsanitize=address,undefined` and the sanitizer runtimes preloaded into the interpreter, so every global, workspace and shared-memory
access of every emulated kernel -- ragged last tiles, padded weight buffers, exactly-sized outputs and dynamic shared memory -- is
bounds-checked, every 8- / 16-byte vector access is checked for CUDA's natural alignment (the emulation's float4 / uint4 carry
alignas(16)), and shifts / signed overflow are checked.  A negative control proves the check is live: the same LayerNorm kernel with an
output buffer one row short must be reported."""

import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _asan():
    libs = []
    for name in ('libasan.so', 'libubsan.so'):
        try:
            path = subprocess.run(['gcc', f'-print-file-name={name}'], capture_output=True, text=True).stdout.strip()
        except OSError:
            return None
        if not (os.path.isabs(path) and os.path.exists(path)):
            return None
        libs.append(path)
    return ' '.join(libs)


ASAN = _asan()
pytestmark = pytest.mark.skipif(ASAN is None, reason='libasan / libubsan not available')


def _env():
    return dict(os.environ, LD_PRELOAD=ASAN, ASAN_OPTIONS='detect_leaks=0:detect_stack_use_after_return=0', ROVITKAN_EMU_SANITIZE='1')


def test_emulated_kernels_are_clean_under_address_and_ub_sanitizers():
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_kernels_on_host.py'), '-q', '-x', '-p', 'no:cacheprovider'],
                       cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=1500)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and ' passed' in out and 'AddressSanitizer' not in out and 'runtime error' not in out, out[-4000:]


def test_address_sanitizer_reports_a_kernel_that_writes_out_of_bounds(tmp_path):
    code = textwrap.dedent(f'''
        import ctypes, pathlib, sys
        import numpy as np
        sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import test_kernels_on_host as T
        k = T.read('encoder_kernels.cu')
        body = ('namespace {{\\n' + T.common_bits()
                + T.between(k, 'constexpr int kD = 192;', '// ------------------------------------------------------------------ patch extraction')
                + T.between(k, '// ------------------------------------------------------------------ LayerNorm forward',
                            '// ------------------------------------------------------------------ LayerNorm backward') + '}}\\n'
                + 'extern "C" void ln(const float* x, const float* g, const float* b, float* y, int rows) {{ EmuDim gr; gr.x = 2; EmuDim bl; bl.x = 128; '
                  'emu_launch(gr, bl, 0, [=] {{ layernorm_fwd_kernel<false, false>(x, 192, g, b, 1e-6f, y, 192, nullptr, nullptr, rows); }}); }}\\n')
        lib = T.compile_host(pathlib.Path({str(tmp_path)!r}), 'oob', body)
        lib.ln.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
        rows = 9
        x, g, b = np.ones((rows, 192), np.float32), np.ones(192, np.float32), np.zeros(192, np.float32)
        y = np.zeros((rows - 1, 192), np.float32)                      # one row short
        lib.ln(T.vp(x), T.vp(g), T.vp(b), T.vp(y), rows)
        print('NOT DETECTED')
    ''')
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and 'AddressSanitizer' in r.stderr and 'heap-buffer-overflow' in r.stderr, (r.stdout + r.stderr)[-3000:]
    assert 'NOT DETECTED' not in r.stdout
