"""TEST INFRASTRUCTURE ONLY -- build/load the g++ SIMT-emulator build of the product's
.cu sources (tests/emul/cuda_emul.h) so kernel *logic* can be checked without a GPU.
The product never loads this library."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libdqa_emul.so")
SRC = os.path.join(ROOT, "perceptron_dqa_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "cuda_emul.h")]
    if not force and os.path.isfile(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-DDQA_EMUL", "-x", "c++", "-I", HERE, "-I", SRC,
           os.path.join(SRC, "dqa_abi.cu"), "-o", SO]
    subprocess.run(cmd, check=True)
    return SO


def load():
    import sys
    sys.path.insert(0, ROOT)
    from perceptron_dqa_b200 import _lib
    return _lib.declare(C.CDLL(build()))


def numpy_workspace(nbytes, _device):
    buf = np.zeros(nbytes + 256, dtype=np.uint8)
    off = (-buf.ctypes.data) % 256
    return buf, buf.ctypes.data + off
