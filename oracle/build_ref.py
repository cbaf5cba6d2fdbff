"""Compiles the reference's OWN native ops from the sources where they lie under /root/reference into
oracle/_ref/ (ORACLE / TEST INFRASTRUCTURE -- the product never loads these).

    python oracle/build_ref.py [--force]

Nothing is copied or patched: g++ / nvcc are pointed at the reference files directly.  Outputs:

  oracle/_ref/iou3d_nms_cuda.so   pcdet/ops/iou3d_nms/src/{iou3d_cpu,iou3d_nms,iou3d_nms_api}.cpp + iou3d_nms_kernel.cu
                                  -> the reference's pybind module: boxes_iou_bev_cpu (runs on CPU here),
                                  nms_gpu, nms_normal_gpu, boxes_overlap_bev_gpu, boxes_iou_bev_gpu (GPU box)
  oracle/_ref/sort_vertices.so    pcdet/ops/rotated_iou/cuda_op/{sort_vert.cpp,sort_vert_kernel.cu}
  oracle/_ref/libref_knn.so       pcdet/ops/knn/src/knn_cuda.cu + oracle/ref_knn_shim.cpp.  knn.cpp itself cannot be
                                  compiled against torch >= 1.11 (knn.cpp:6,9 include the removed THC headers), so
                                  the kernel launcher it wraps (knn_cuda.cu:97) is called through a 10-line extern "C" shim.

-O2 is required: iou3d_cpu.cpp and iou3d_nms_kernel.cu define the same inline functions, and at -O0 the
linker resolves the CPU call to nvcc's host stub of the __device__ function.
The reference's own build system (setup.py / setuptools) is NOT run.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("CG3D_REFERENCE", "/root/reference")
OPS = os.path.join(REF, "pcdet", "ops")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _torch_flags():
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", "-I/usr/local/cuda/include"]
    libdirs = ce.library_paths()
    link = [f"-L{p}" for p in libdirs] + [f"-Wl,-rpath,{p}" for p in libdirs] + \
           ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart"]
    return inc, link


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-4000:])
        raise RuntimeError("reference build step failed")


def _obj(src, name, inc, extra):
    o = os.path.join(OUT, "obj", name + ".o")
    if os.path.exists(o) and os.path.getmtime(o) > os.path.getmtime(src):
        return o
    common = ["-O2", "-std=c++17", "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1", *extra, *inc]
    if src.endswith(".cu"):
        _run([NVCC, *ARCH, *common, "-Xcompiler", "-fPIC", "-w", "-c", src, "-o", o])
    else:
        _run(["g++", *common, "-fPIC", "-w", "-c", src, "-o", o])
    return o


def build(force: bool = False) -> bool:
    if not os.path.isdir(OPS):
        return False
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    targets = {
        "iou3d_nms_cuda.so": ("iou3d_nms_cuda", [os.path.join(OPS, "iou3d_nms", "src", f) for f in
                                                 ("iou3d_cpu.cpp", "iou3d_nms.cpp", "iou3d_nms_api.cpp", "iou3d_nms_kernel.cu")]),
        "sort_vertices.so": ("sort_vertices", [os.path.join(OPS, "rotated_iou", "cuda_op", f) for f in
                                               ("sort_vert.cpp", "sort_vert_kernel.cu")]),
        "libref_knn.so": (None, [os.path.join(OPS, "knn", "src", "knn_cuda.cu"), os.path.join(HERE, "ref_knn_shim.cpp")]),
    }
    inc, link = _torch_flags()
    jobs = []
    for so, (mod, srcs) in targets.items():
        path = os.path.join(OUT, so)
        if not force and os.path.exists(path) and all(os.path.getmtime(path) > os.path.getmtime(s) for s in srcs):
            continue
        extra = [f"-DTORCH_EXTENSION_NAME={mod}"] if mod else []
        jobs.append((path, mod, srcs, extra))
    with ThreadPoolExecutor(max_workers=8) as ex:
        for path, mod, srcs, extra in jobs:
            tag = os.path.basename(path).split(".")[0]
            objs = list(ex.map(lambda s: _obj(s, tag + "_" + os.path.basename(s).replace(".", "_"),
                                              inc if mod else ["-I/usr/local/cuda/include"], extra), srcs))
            _run(["g++", "-shared", "-o", path, *objs, *(link if mod else ["-L/usr/local/cuda/lib64", "-lcudart"])])
    return True


def load(name: str):
    """import a built reference pybind module (iou3d_nms_cuda / sort_vertices) from oracle/_ref."""
    import importlib.util
    import torch  # noqa: F401  (the modules link against libtorch)
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("reference ops built into", OUT if ok else "(skipped: /root/reference absent)")
