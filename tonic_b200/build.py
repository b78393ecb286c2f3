"""Builds the sm_100a shared library `tonic_b200/libtonic_b200.so` in-tree with
nvcc (cross-compiles without a GPU).  Used by `__graft_entry__.build()`; the
built library travels to the GPU box with the repo snapshot."""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libtonic_b200.so')
SOURCES = ['api.cu', 'env_step.cu', 'classic_env.cu', 'returns.cu', 'moments.cu', 'mlp.cu', 'optim.cu',
           'heads.cu', 'offpolicy.cu', 'tc_gemm.cu', 'tc_mlp.cu', 'host_rng.cpp']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def newest_source_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(os.path.dirname(HERE), 'include', 'tonic_b200.h'))
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest_source_mtime():
        return LIB
    nvcc = find_nvcc()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    def compile_one(src):
        obj = os.path.join(objdir, src.rsplit('.', 1)[0] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        # an object newer than its own source and every header is reused
        deps = [os.path.join(CSRC, src), os.path.join(os.path.dirname(HERE), 'include', 'tonic_b200.h')]
        deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(map(os.path.getmtime, deps)):
            return obj, f'$ (up to date) {obj}\n', 0
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, f'$ {" ".join(cmd)}\n{res.stdout}{res.stderr}', res.returncode

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    objects = [r[0] for r in results]
    logs = [r[1] for r in results]
    for obj, log, rc in results:
        if rc != 0:
            raise RuntimeError('nvcc failed:\n' + log)
    cmd = [nvcc, '-shared', '-o', LIB] + objects + ['-lcudart']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    with open(os.path.join(objdir, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(logs))
    if verbose:
        print('\n'.join(logs))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
