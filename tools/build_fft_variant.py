"""Build a variant of libpmgrav.so that differs only in pm_fft.cu's -D flags (development aid; seconds instead of minutes):

    python tools/build_fft_variant.py <name> [-DPM_FFT_CY_BYTES=128 ...]   ->  concept_b200/lib/libpmgrav_<name>.so

The other objects are the ones `python -m concept_b200.build` left in concept_b200/lib/."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concept_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    nvcc = '/usr/local/cuda/bin/nvcc'
    nccl_inc, nccl_lib = B._nccl_paths()
    cuda_lib = '/usr/local/cuda/lib64'
    obj = os.path.join(B.LIBDIR, f'pm_fft_{name}.var.o')
    subprocess.run([nvcc, *B.ARCH, *B.NVCC_FLAGS, *flags, '-I', os.path.join(ROOT, 'include'), '-I', nccl_inc, '-c',
                    os.path.join(B.CSRC, 'pm_fft.cu'), '-o', obj], check=True)
    objs = [o for o in glob.glob(os.path.join(B.LIBDIR, 'pm_*.o')) if not o.endswith('.var.o') and os.path.basename(o) != 'pm_fft.o']
    lib = os.path.join(B.LIBDIR, f'libpmgrav_{name}.so')
    subprocess.run([nvcc, *B.ARCH, '-shared', '-o', lib, obj, *objs, '-L', cuda_lib, '-L', nccl_lib, '-lcufft', '-l:libnccl.so.2',
                    '-Xlinker', f'-rpath={nccl_lib}', '-Xlinker', f'-rpath={cuda_lib}', '-Xlinker', '-rpath=/usr/lib/x86_64-linux-gnu'],
                   check=True)
    os.remove(obj)
    print(lib)


if __name__ == '__main__':
    main()
