"""Builds univer_ocr_b200/lib/libuocr.so from csrc/*.cu with nvcc for sm_100a (in-tree, so the
library travels to the GPU box with the repository snapshot).

    python -m univer_ocr_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
OBJ_DIR = os.path.join(PKG_DIR, 'build')
LIB_DIR = os.path.join(PKG_DIR, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libuocr.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
    '-I', INCLUDE, '-I', CSRC,
]


def find_nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: libuocr.so cannot be built')
    return nvcc


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    return hs + [os.path.join(INCLUDE, 'uocr.h')]


def build(force=False, verbose=False, extra_flags=()):
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    hdr_time = _newest(headers() + [os.path.abspath(__file__)])
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        stale = (force or not os.path.exists(obj)
                 or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time))
        jobs.append((src, obj, stale))

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return obj, ''
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, '-c', src, '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{res.stdout}\n{res.stderr}')
        return obj, res.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
        results = list(pool.map(compile_one, jobs))
    if verbose:
        for obj, log in results:
            if log.strip():
                print(f'--- {os.path.basename(obj)}\n{log}')
    objs = [o for o, _ in results]
    if force or any(j[2] for j in jobs) or not os.path.exists(LIB_PATH):
        cmd = [nvcc, '-shared', '-o', LIB_PATH, *objs, '-lcudart', '-ldl', '-Xlinker', '--no-undefined']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    return LIB_PATH


if __name__ == '__main__':
    flags = ('-Xptxas', '-v') if '--verbose' in sys.argv else ()
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, extra_flags=flags)
    print(path)
