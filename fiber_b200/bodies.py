"""Out-of-tree device bodies: compile a CUDA translation unit into a body module (shared object).

The reference pickles any callable into the task tuple (fiber/pool.py:961).  On this engine a callable
needs device code; ``compile_module`` turns user CUDA source (written against
``include/fiber_b200_body.cuh``) into a shared object that ``fbr_register_body`` loads.  Modules are
cached by content hash under ``fiber_b200/_lib/bodies/`` -- in-tree, so a module built on a machine
without a GPU travels to the GPU box with the rest of the build.
"""
import hashlib
import os
import subprocess

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MODDIR = os.path.join(_build.LIBDIR, "bodies")
INCLUDES = [os.path.join(ROOT, "include"), _build.CSRC]


def _stamp():
    """Hash of the headers a module is compiled against: a header change rebuilds every module."""
    h = hashlib.blake2b(digest_size=8)
    for path in (os.path.join(ROOT, "include", "fiber_b200.h"), os.path.join(ROOT, "include", "fiber_b200_body.cuh"),
                 os.path.join(_build.CSRC, "kernels.cuh"), os.path.join(_build.CSRC, "bodies.cuh")):
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.digest()


def module_path(name, source):
    key = hashlib.blake2b(source.encode() + _stamp(), digest_size=8).hexdigest()
    return os.path.join(MODDIR, "%s-%s.so" % (name, key))


def compile_module(name, source, force=False):
    """nvcc the translation unit ``source`` for sm_100a into a body module; returns its path."""
    so = module_path(name, source)
    if os.path.exists(so) and not force:
        return so
    os.makedirs(MODDIR, exist_ok=True)
    cu = so[:-3] + ".cu"
    with open(cu, "w") as fh:
        fh.write(source)
    cmd = [_build.nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared"] + ["-I" + d for d in INCLUDES] + [cu, "-o", so + ".tmp"]
    try:
        subprocess.check_output(cmd, stderr=subprocess.STDOUT)
    except subprocess.CalledProcessError as e:
        raise RuntimeError("nvcc failed for device body %r:\n%s" % (name, e.output.decode("utf-8", "replace"))) from None
    os.replace(so + ".tmp", so)
    return so
