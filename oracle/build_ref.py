"""TEST INFRASTRUCTURE ONLY.  Recipe that compiles the reference's own native
sources for the hot path, from where they lie under /root/reference, into
``oracle/_ref/`` (git-ignored; travels to the GPU box with the snapshot).

  oracle/_ref/cython_nms*.so      <- lib/utils/cython_nms.pyx   (CPU NMS, the parity target)
  oracle/_ref/cython_bbox*.so     <- lib/utils/cython_bbox.pyx  (needed to import utils.boxes)
  oracle/_ref/libref_roialign.so  <- lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu
                                     compiled UNMODIFIED for sm_100a (GPU oracle + speed baseline)
  oracle/_ref/libref_flowalign.so <- lib_vos/vos_model/flow_align/src/flow_align_cuda_kernel.cu
                                     compiled UNMODIFIED for sm_100a (GPU oracle + speed baseline)

No reference source is copied into the repo: the .pyx is read, a 2-token
NumPy-2 compatibility substitution (``np.int_t``->``np.int64_t``,
``dtype=np.int``->``dtype=np.int64``; lib/utils/cython_nms.pyx:45,48,49) is
applied to the in-memory text, and the cythonized C lives in a temp dir.
The reference's shipped cython_nms.c (Cython 0.29.1) does not compile on
Python 3.12, and its build system (lib/setup.py, lib/make.sh) is not used.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("VOSD_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def _cython_ext(name, text, tmp):
    import numpy as np
    pyx = os.path.join(tmp, name + ".pyx")
    with open(pyx, "w") as f:
        f.write(text)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", os.path.join(tmp, name + ".c")])
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(OUT, name + suffix)
    # -O2, no -march=native: keeps the reference's baseline x86-64 (no-FMA) arithmetic
    subprocess.check_call([
        "gcc", "-shared", "-fPIC", "-O2", "-Wno-cpp", "-Wno-unused-function",
        "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
        os.path.join(tmp, name + ".c"), "-o", out])
    return out


def build(force=False):
    if not os.path.isdir(os.path.join(REF_ROOT, "lib")):
        print("[build_ref] %s absent: keeping prebuilt oracle/_ref as is" % REF_ROOT)
        return False
    os.makedirs(OUT, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    tmp = tempfile.mkdtemp(prefix="vosd_ref_")
    try:
        if force or not os.path.exists(os.path.join(OUT, "cython_nms" + suffix)):
            text = open(os.path.join(REF_ROOT, "lib/utils/cython_nms.pyx")).read()
            text = text.replace("np.int_t", "np.int64_t").replace("dtype=np.int)", "dtype=np.int64)")
            _cython_ext("cython_nms", text, tmp)
        if force or not os.path.exists(os.path.join(OUT, "cython_bbox" + suffix)):
            text = open(os.path.join(REF_ROOT, "lib/utils/cython_bbox.pyx")).read()
            _cython_ext("cython_bbox", text, tmp)
        so = os.path.join(OUT, "libref_roialign.so")
        if force or not os.path.exists(so):
            src = os.path.join(REF_ROOT, "lib/modeling/roi_xfrom/roi_align/src")
            subprocess.check_call([
                "nvcc", "-shared", "-Xcompiler", "-fPIC", "-O3",
                "-gencode", "arch=compute_100a,code=sm_100a", "-I" + src,
                "-x", "cu", os.path.join(src, "roi_align_kernel.cu"), "-o", so])
        so = os.path.join(OUT, "libref_flowalign.so")
        if force or not os.path.exists(so):
            src = os.path.join(REF_ROOT, "lib_vos/vos_model/flow_align/src")
            subprocess.check_call([
                "nvcc", "-shared", "-Xcompiler", "-fPIC", "-O3",
                "-gencode", "arch=compute_100a,code=sm_100a", "-I" + src,
                "-x", "cu", os.path.join(src, "flow_align_cuda_kernel.cu"), "-o", so])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(sorted(os.listdir(OUT)) if os.path.isdir(OUT) else "no _ref")
