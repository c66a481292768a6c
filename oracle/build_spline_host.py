"""TEST INFRASTRUCTURE ONLY: compile oracle/spline_host.cpp (which includes the product header nnb_spline.cuh as plain C++)
into oracle/_build/libspline_host.so with g++.  No GPU, no nvcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_build', 'libspline_host.so')


def build():
    src = os.path.join(HERE, 'spline_host.cpp')
    hdr = os.path.join(HERE, '..', 'nnest_b200', 'csrc', 'nnb_spline.cuh')
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off', '-x', 'c++', src, '-o', OUT])
    return OUT


if __name__ == '__main__':
    print(build())
