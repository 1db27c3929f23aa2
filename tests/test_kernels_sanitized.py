"""The emulated kernels under AddressSanitizer + UndefinedBehaviorSanitizer (the CPU stand-in for `compute-sanitizer --tool memcheck`,
which needs a GPU): the host-emulation suite (tests/test_kernels_on_host.py) is re-run in a subprocess with the harnesses compiled
`-fsanitize=address,undefined` and the sanitizer runtimes preloaded into the interpreter, so every global, workspace and shared-memory
access of every emulated kernel -- ragged last tiles, padded weight buffers, exactly-sized outputs and dynamic shared memory -- is
bounds-checked, every 8- / 16-byte vector access is checked for CUDA's natural alignment (the emulation's float4 / uint4 carry
alignas(16)), and shifts / signed overflow are checked.  A negative control proves the check is live: the same LayerNorm kernel with an
output buffer one row short must be reported.

The sanitized run also resumes the fibers in REVERSE thread order (ROVITKAN_EMU_ORDER=reverse; the plain run uses ascending order):
a kernel with a missing __syncthreads / __syncwarp, whose result depends on which thread reaches a barrier-free stretch first, cannot
pass under both orders -- the emulation's stand-in for `compute-sanitizer --tool racecheck`."""

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
    return dict(os.environ, LD_PRELOAD=ASAN, ASAN_OPTIONS='detect_leaks=0:detect_stack_use_after_return=0', ROVITKAN_EMU_SANITIZE='1', ROVITKAN_EMU_ORDER='reverse')


def test_emulated_kernels_are_clean_under_address_and_ub_sanitizers():
    # (the three-layer chain and the second forward-GEMM test re-run kernels that other tests of the pass already cover: left out to
    # keep the suite short)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_kernels_on_host.py'), '-q', '-x', '-p', 'no:cacheprovider',
                        '-k', 'not kan_stack and not fused_epilogues'],
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


def test_the_two_fiber_orders_expose_a_missing_barrier(tmp_path):
    """Control for the order check: a kernel that reads its neighbour's shared-memory slot WITHOUT a __syncthreads gives different
    results under the ascending and the reverse fiber order; with the barrier both orders agree."""
    code = textwrap.dedent(f'''
        import ctypes, pathlib, sys
        import numpy as np
        sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import test_kernels_on_host as T
        body = """
        template <bool SYNC> __global__ void neighbour(float* out, float base) {{
          __shared__ float s[64];
          s[threadIdx.x] = base + threadIdx.x;
          if (SYNC) __syncthreads();
          out[threadIdx.x] = s[(threadIdx.x + 1) % 64];
        }}
        extern "C" void run(int sync, float* out) {{
          EmuDim g; EmuDim b; b.x = 64;
          if (sync) emu_launch(g, b, 0, [=] {{ neighbour<true>(out, 100.0f); }}); else emu_launch(g, b, 0, [=] {{ neighbour<false>(out, 100.0f); }});
          if (sync) emu_launch(g, b, 0, [=] {{ neighbour<true>(out + 64, 1.0f); }}); else emu_launch(g, b, 0, [=] {{ neighbour<false>(out + 64, 1.0f); }});
        }}
        """
        lib = T.compile_host(pathlib.Path({str(tmp_path)!r}), 'racy_' + sys.argv[1], body)
        lib.run.argtypes = [ctypes.c_int, ctypes.c_void_p]
        for sync in (0, 1):
            out = np.zeros(128, np.float32)
            lib.run(sync, T.vp(out))
            print('RESULT', sync, out[64:].tolist())       # second launch: the static shared array holds the previous block's values
    ''')
    res = {}
    for order in ('ascending', 'reverse'):
        env = dict(os.environ, ROVITKAN_EMU_ORDER=order)
        env.pop('ROVITKAN_EMU_SANITIZE', None)
        r = subprocess.run([sys.executable, '-c', code, order], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[order] = {l.split()[1]: l.split(None, 2)[2] for l in r.stdout.splitlines() if l.startswith('RESULT')}
    want = str([float(1 + (t + 1) % 64) for t in range(64)])
    assert res['ascending']['1'] == res['reverse']['1'] == want
    assert res['ascending']['0'] != res['reverse']['0'] and res['ascending']['0'] != want and res['reverse']['0'] != want


def test_the_emulation_reports_a_divergent_barrier_instead_of_hanging(tmp_path):
    """Control for the watchdog: a kernel in which part of a block waits at a __syncthreads and the rest at a full-warp shuffle that
    half of its warp never executes (threads that RETURN are dropped from the barriers, as on the GPU; threads that stay alive and
    never arrive are a deadlock) must abort with a message, not spin forever."""
    code = textwrap.dedent(f'''
        import ctypes, pathlib, sys
        sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import test_kernels_on_host as T
        body = """
        __global__ void divergent(float* out) {{
          float v = 1.0f + threadIdx.x;
          if (threadIdx.x < 32 || (threadIdx.x & 16)) __syncthreads();   // 48 of the 64 threads arrive at the block barrier ...
          else v = __shfl_xor_sync(0xffffffffu, v, 1);                   // ... 16 lanes of a 32-lane warp wait for a full-warp shuffle
          out[threadIdx.x] = v;
        }}
        extern "C" void run(float* out) {{ EmuDim g; EmuDim b; b.x = 64; emu_launch(g, b, 0, [=] {{ divergent(out); }}); }}
        """
        lib = T.compile_host(pathlib.Path({str(tmp_path)!r}), 'dead', body)
        import numpy as np
        out = np.zeros(64, np.float32)
        lib.run.argtypes = [ctypes.c_void_p]
        lib.run(T.vp(out))
        print('RETURNED')
    ''')
    env = dict(os.environ)
    env.pop('ROVITKAN_EMU_SANITIZE', None)
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and 'cuda_host_emu: deadlock' in r.stderr and 'RETURNED' not in r.stdout, (r.stdout + r.stderr)[-2000:]
